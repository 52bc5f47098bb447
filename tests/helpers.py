"""Shared helpers for the parity tests: small versions of the BASELINE configs, hand-made triangle
scenes on an identity camera, and comparison utilities."""
import ctypes as C
import math

import numpy as np

import swraster_viewer_b200 as swr
from swraster_viewer_b200 import abi, scenes

SMALL = dict(voxel_dim=8, cube_size=32)


def small_configs():
    """(name, scene, camera spec, W, H): scaled-down C1..C5, the all-materials scene and the translucent scene."""
    return [
        ("c1_sphere", *scenes.scene_c1_sphere(48, 48, **SMALL), 448, 256),
        ("c2_terrain", *scenes.scene_c2_terrain(96, 128, **SMALL), 448, 256),
        ("c3_instanced", *scenes.scene_c3_instanced(60000, ico_subdiv=2, torus_n=10, box_n=3, **SMALL), 512, 288),
        ("c4_micro", *scenes.scene_c4_micro(161, **SMALL), 256, 144),
        ("c5_shards", *scenes.scene_c5_shards(4, 65, **SMALL), 384, 216),
        ("materials", *scenes.scene_materials_test(**SMALL), 448, 256),
        ("translucent", *scenes.scene_translucent_test(**SMALL), 448, 256),
    ]


def identity_camera(W, H, intersecting=False):
    """clip = position (view_project = identity). All primitives classify Inside, or Intersecting (-> clipper)."""
    cam = abi.Camera()
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    cam.view_matrix = (C.c_float * 16)(*ident)
    cam.view_project_matrix = (C.c_float * 16)(*ident)
    cam.skybox_matrix_transposed = (C.c_float * 16)(*ident)
    for p in range(6):
        cam.view_clip_planes[p] = (C.c_float * 4)(0, 0, 0, 0.0 if intersecting else 1e30)
    cam.position = (C.c_float * 4)(0, 0, 5, 0)
    cam.one_over_width, cam.one_over_height = 1.0 / W, 1.0 / H

    class Cam:
        pass
    c = Cam()
    c.abi, c.width, c.height = cam, W, H
    return c


def ndc_from_subpixel(X, Y, W, H):
    """Clip-space x,y (w=1) whose snap (renderer.rs:834-846) lands exactly on sub-pixel (X, Y)."""
    return 2.0 * (X / 16.0) / W - 1.0, 1.0 - 2.0 * (Y / 16.0) / H


def triangle_scene(tris_subpx, W, H, depths=None, **kw):
    """One primitive, identity node. tris_subpx: list of 3x(X,Y) sub-pixel corners (screen space, y down).
    depths: per-triangle list of 3 z values (clip z with w=1)."""
    pos, idx = [], []
    for t, tri in enumerate(tris_subpx):
        for k, (X, Y) in enumerate(tri):
            x, y = ndc_from_subpixel(X, Y, W, H)
            z = 0.5 if depths is None else depths[t][k]
            pos.append([x, y, z])
            idx.append(len(pos) - 1)
    pos = np.array(pos, np.float64)
    n = len(pos)
    nrm = np.tile(np.array([[0.0, 0.0, 1.0]]), (n, 1))
    uv = (pos[:, :2] * 0.5 + 0.5)
    prim = scenes._finish_prim(pos, nrm, uv, np.array(idx), 0)
    kw = {**SMALL, **kw}
    return scenes.SceneData([[prim]], [scenes.Node(scenes.IDENT, 0)], [scenes.Material((0.8, 0.6, 0.4, 1), 0.2, 0.6)], [], **kw)


def rgba_bytes(px):
    return np.stack([(px >> 24) & 255, (px >> 16) & 255, (px >> 8) & 255], -1).astype(np.int32)


def render_gpu(scene, cam, W, H, device=0, rows=None, reference_rsqrt=True):
    r = swr.Renderer(W, H, device)
    r.set_reference_rsqrt(reference_rsqrt)
    if rows is not None:
        r.set_tile_rows(*rows)
    r.render_scene(scene, cam)
    buf = swr.RenderBuffer(W, H)
    r.blit_to_buffer(buf)
    depth, seq, b1, b2 = r.read_visbuffer()
    out = dict(depth=depth, seq=seq, bary1=b1, bary2=b2, color=r.read_color(), pixels=buf.pixels.copy(),
               luminance=r.read_tile_luminance(), stats=r.stats(), rsqrt_bits=r.reference_rsqrt_bits)
    r.close()
    return out


def render_oracle(scene, cam, W, H, exact_rsqrt=False):
    import oracle as orc
    orc.set_exact_rsqrt(exact_rsqrt)
    o = orc.Oracle(W, H)
    out = o.render(scene, cam.abi, nthreads=1)
    out["pixels"] = o.resolve(2.0)
    out["stats"] = o.stats.as_dict()
    orc.set_exact_rsqrt(False)
    return out

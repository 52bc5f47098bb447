"""The CHECKER of the device sun-visibility bake (oracle/oracle_sunvis.cpp) against a brute-force float32
restatement of gi.rs:151-314 / raytracer.rs:177-259 / voxelgrid.rs:371-419 written here in numpy. CPU only: this pins the
oracle that tests/test_gpu_bakes.py then holds the CUDA ray cast (csrc/swr_bake.cuh) to, end to end included."""
import math

import numpy as np
import pytest

import oracle as orc
from swraster_viewer_b200 import abi, scenes

F32 = np.float32


def shadow_scene(translucent_lid=False, voxel_dim=10):
    """A ground slab, a floating box above it and (optionally) a translucent sheet above half of the ground."""
    mats = [scenes.Material((0.8, 0.8, 0.8, 1), 0.0, 0.8), scenes.Material((0.9, 0.2, 0.2, 1), 0.0, 0.5)]
    if translucent_lid:
        mats.append(scenes.Material((0.2, 0.4, 0.9, 1), 0.0, 0.3, flags=abi.MAT_TRANSLUCENT, transmission=0.5))
    meshes = [[scenes.box_grid(2, 0)], [scenes.box_grid(2, 1)]]
    nodes = [scenes.Node(scaled((0, -0.25, 0), (6.0, 0.25, 6.0)), 0), scenes.Node(scaled((0.5, 3.0, -0.5), (0.9, 0.3, 0.9)), 1)]
    if translucent_lid:
        meshes.append([scenes.box_grid(1, 2)])
        nodes.append(scenes.Node(scaled((-3.0, 4.0, 0.0), (2.5, 0.02, 5.0)), 2))
    sc = scenes.SceneData(meshes, nodes, mats, [], voxel_dim=voxel_dim, cube_size=16, seed=4)
    cam = scenes.CameraSpec((0.0, 9.0, 9.0), (0.0, 0.0, 0.0), math.pi / 4, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


def scaled(t, s):
    m = np.eye(4)
    m[0, 0], m[1, 1], m[2, 2] = s
    m[:3, 3] = t
    return np.ascontiguousarray(m.T.reshape(-1).astype(np.float32))


def f32(x):
    return np.asarray(x, F32)


def brute_force(sc):
    d = sc.desc()
    W, H, D = d.voxel_grid.dims[:]
    wmin, wmax = f32(d.voxel_grid.world_min[:]), f32(d.voxel_grid.world_max[:])
    vs = (wmax - wmin) / f32([W, H, D])
    tris, mat = [], []
    for n in sc.nodes:
        M = n.transform.reshape(4, 4).T.astype(F32)
        for p in sc.meshes[n.mesh_index]:
            P = np.stack([((M[:, 0] * v[0] + M[:, 1] * v[1]) + M[:, 2] * v[2]) + M[:, 3] * v[3] for v in p.positions.astype(F32)])[:, :3]
            for t in range(0, len(p.indices), 3):
                a, b, c = (P[p.indices[t + k]] for k in range(3))
                if f32(np.dot(np.cross(b - a, c - a), np.cross(b - a, c - a))) <= 1e-12:
                    continue
                tris.append((a, b, c))
                mat.append(p.material_index)
    tris = f32(tris)
    p0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    # active mask (gi.rs:151-265)
    occ = np.zeros((D, H, W), bool)
    area_ref = max(min(vs), 1e-6) ** 2
    for (a, b, c) in tris:
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a))
        target = int(min(max(math.ceil(f32(area / area_ref * 2.0)), 1), 4096))
        n = int(math.ceil(math.sqrt(target)))
        for iu in range(n):
            for iv in range(n - iu):
                u, v = F32((iu + 0.5) / n), F32((iv + 0.5) / n)
                w = F32(F32(1.0) - u) - v
                if w < 0:
                    continue
                pt = (a * w + b * u) + c * v
                f = (pt - wmin) / vs
                if (f < 0).any():
                    continue
                ix = np.floor(f).astype(int)
                if ix[0] >= W or ix[1] >= H or ix[2] >= D:
                    continue
                occ[ix[2], ix[1], ix[0]] = True
    act = np.zeros_like(occ)
    for z, y, x in zip(*np.nonzero(occ)):
        act[max(z - 2, 0):z + 3, max(y - 2, 0):y + 3, max(x - 2, 0):x + 3] = True
    L = f32(d.light_direction[:])
    L = L / F32(np.sqrt(F32(F32(L[0] * L[0] + L[1] * L[1]) + L[2] * L[2])))
    bias = F32(np.linalg.norm(vs.astype(np.float64))) * F32(3.0)
    out = np.ones((D, H, W), F32)
    pvec = np.cross(L[None], e2).astype(F32)
    det = np.einsum("ij,ij->i", e1, pvec).astype(F32)
    ok = np.abs(det) > 1e-8
    inv = np.where(ok, F32(1.0) / np.where(ok, det, 1), 0).astype(F32)
    for z, y, x in zip(*np.nonzero(act)):
        o = (wmin + vs * F32(0.5)) + f32([x, y, z]) * vs + L * bias
        tv = (o[None] - p0).astype(F32)
        u = np.einsum("ij,ij->i", tv, pvec).astype(F32) * inv
        q = np.cross(tv, e1).astype(F32)
        v = (q @ L).astype(F32) * inv
        t = np.einsum("ij,ij->i", e2, q).astype(F32) * inv
        hit = ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t >= 1e-4)
        tr = F32(1.0)
        for k in np.argsort(np.where(hit, t, np.inf), kind="stable")[:int(hit.sum())]:
            m = sc.materials[mat[k]]
            if not (m.flags & abi.MAT_TRANSLUCENT):
                tr = F32(0.0)
                break
            tr = F32(tr * F32(m.transmission))
            if tr <= 1e-4:
                tr = F32(0.0)
                break
        out[z, y, x] = tr
    if not act.any():
        return out, act
    from scipy.ndimage import uniform_filter  # 3x3x3 mean over the neighbours that lie inside the grid, squared
    num = uniform_filter(out.astype(np.float64), 3, mode="constant", cval=0.0)
    den = uniform_filter(np.ones_like(out, np.float64), 3, mode="constant", cval=0.0)
    blur = (F32(1.0) * (num / den).astype(F32)) ** 2
    return blur, act


@pytest.mark.parametrize("lid", [False, True])
def test_sun_visibility_matches_the_brute_force_restatement(lid):
    sc, _ = shadow_scene(lid, voxel_dim=32)
    got = orc.sun_visibility(sc)
    want, active = brute_force(sc)
    assert np.array_equal(orc.sunvis_active_mask(sc).astype(bool), active)
    assert active.any() and not active.all() and got.shape == (32, 32, 32)
    assert np.allclose(got, want, rtol=0, atol=1e-5), np.abs(got - want).max()
    assert got.min() < 0.05 and got.max() == 1.0  # some voxels sit in the box's shadow, some see the sun
    if lid:
        mid = (got > 0.15) & (got < 0.35)  # under the 0.5-transmission sheet: 0.5, squared by the blur = 0.25
        assert mid.any()


def test_empty_scene_keeps_every_voxel_lit():
    sc, _ = shadow_scene(False, voxel_dim=4)
    sc.nodes = []
    sc._desc = None
    sc.node_spheres = []
    assert np.array_equal(orc.sun_visibility(sc), np.ones((4, 4, 4), F32))  # no active voxel: no rays, no blur (gi.rs:281-284)

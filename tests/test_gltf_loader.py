"""Native glTF / GLB loader (include/swr_gltf.h) against (a) the in-memory procedural scenes it must
reproduce after an export -> load round trip, (b) independent restatements of the reference's loader arithmetic written
here in plain Python / numpy (scene.rs:504-646, texture.rs:45-128), (c) hand-built documents for the container rules of the
glTF 2.0 specification (GLB, data URIs, strides, sparse accessors, normalised integers), (d) the reference's error wording.
CPU only: nothing here touches a device."""
import base64
import io
import json
import struct
import zlib

import numpy as np
import pytest

from swraster_viewer_b200 import abi, gltf, scenes
from helpers import SMALL, render_oracle
import swraster_viewer_b200 as swr

F32 = np.float32


# ---------------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------------
def prim_arrays(d, i):
    p = d.primitives[i]
    n = p.nverts
    pos = np.ctypeslib.as_array(p.positions, (n, 4)).copy()
    nrm = np.ctypeslib.as_array(p.normals, (n, 4)).copy()
    tan = np.ctypeslib.as_array(p.tangents, (n, 4)).copy()
    uv = np.ctypeslib.as_array(p.texcoords, (n, 2)).copy()
    idx = np.ctypeslib.as_array(p.indices, (p.nindices,)).copy()
    return pos, nrm, tan, uv, idx


def tex_arrays(t):
    nm = t.max_mip_level + 1
    return (np.ctypeslib.as_array(t.data, (t.ntexels,)).copy(), np.ctypeslib.as_array(t.mip_offsets, (nm,)).copy(),
            np.ctypeslib.as_array(t.mip_widths, (nm,)).copy(), np.ctypeslib.as_array(t.mip_heights, (nm,)).copy(),
            np.ctypeslib.as_array(t.array_stride, (nm,)).copy())


def channel_diff(a, b):
    sh = np.array([24, 16, 8, 0], np.uint32)
    return np.abs(((a[:, None] >> sh) & 255).astype(np.int32) - ((b[:, None] >> sh) & 255).astype(np.int32))


def write_gltf(tmp_path, doc, name="doc.gltf", blobs=None):
    for fn, data in (blobs or {}).items():
        (tmp_path / fn).write_bytes(data)
    (tmp_path / name).write_text(json.dumps(doc))
    return str(tmp_path / name)


def tri_doc(extra_prim=None, extra_root=None, positions=None, indices=None):
    """One triangle in a single buffer given as a data URI."""
    pos = np.array(positions if positions is not None else [[0, 0, 0], [1, 0, 0], [0, 1, 0]], F32)
    idx = np.array(indices if indices is not None else [0, 1, 2], np.uint16)
    blob = pos.tobytes() + idx.tobytes()
    blob += b"\0" * (-len(blob) % 4)
    doc = {"asset": {"version": "2.0"},
           "buffers": [{"uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode(), "byteLength": len(blob)}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": pos.nbytes}, {"buffer": 0, "byteOffset": pos.nbytes, "byteLength": idx.nbytes}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": len(pos), "type": "VEC3"},
                         {"bufferView": 1, "componentType": 5123, "count": len(idx), "type": "SCALAR"}],
           "materials": [{}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "material": 0}]}],
           "nodes": [{"mesh": 0}]}
    if extra_prim:
        doc["meshes"][0]["primitives"][0].update(extra_prim)
    if extra_root:
        doc.update(extra_root)
    return doc


# ---------------------------------------------------------------------------------------------------------------------
# (a) round trip of the procedural scenes
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("maker", ["scene_materials_test", "scene_translucent_test"])
def test_roundtrip_reproduces_the_scene(tmp_path, maker):
    sc, spec = getattr(scenes, maker)(**SMALL)
    scenes.export_gltf(sc, str(tmp_path / "scene"))
    g = gltf.load_gltf(tmp_path / "scene.gltf", environment=sc)
    d, ref = g.desc(), sc.desc()
    assert (d.nprimitives, d.nmeshes, d.nnodes, d.nmaterials, d.ntextures) == (ref.nprimitives, ref.nmeshes, ref.nnodes, ref.nmaterials, ref.ntextures)
    flat = [p for m in sc.meshes for p in m]
    for i, p in enumerate(flat):
        pos, nrm, tan, uv, idx = prim_arrays(d, i)
        assert np.array_equal(pos.view(np.uint32), p.positions.view(np.uint32))
        assert np.array_equal(nrm[:, :3].view(np.uint32), np.ascontiguousarray(p.normals[:, :3]).view(np.uint32))
        assert np.array_equal(tan.view(np.uint32), p.tangents.view(np.uint32))
        assert np.array_equal(uv.view(np.uint32), p.texcoords.view(np.uint32))
        assert np.array_equal(idx, p.indices)
        assert d.primitives[i].material_index == p.material_index
        assert np.array_equal(np.array(d.primitives[i].bounding_sphere[:], F32).view(np.uint32), np.array(p.bounding_sphere(), F32).view(np.uint32))
    for i in range(d.nmeshes):
        assert (d.meshes[i].first_primitive, d.meshes[i].num_primitives) == (ref.meshes[i].first_primitive, ref.meshes[i].num_primitives)
    for i, n in enumerate(sc.nodes):
        assert np.array_equal(np.array(d.nodes[i].transform[:], F32).view(np.uint32), n.transform.view(np.uint32))
        assert d.nodes[i].mesh_index == n.mesh_index
        # node sphere: centre = node origin (scene.rs:384-386), radius grown over the primitives (scene.rs:318-329)
        assert np.allclose(d.nodes[i].bounding_sphere_world[:], sc.node_spheres[i], rtol=2e-6, atol=1e-6)
    for i, m in enumerate(sc.materials):
        gm = d.materials[i]
        assert np.array_equal(np.array(gm.base_color_factor[:], F32), np.array(m.base_color_factor, F32))
        assert (gm.metallic_factor, gm.roughness_factor) == (F32(m.metallic_factor), F32(m.roughness_factor))
        assert np.array_equal(np.array(gm.emissive_factor[:], F32), np.array(m.emissive_factor, F32))
        assert gm.flags == m.flags and gm.alpha_cutoff == F32(m.alpha_cutoff) and gm.transmission == F32(m.transmission)
        assert gm.occlusion_strength == F32(m.occlusion_strength)
        for k in ("base_color_texture", "metallic_roughness_texture", "normal_texture", "emissive_texture", "occlusion_texture", "transmission_texture"):
            assert getattr(gm, k) == getattr(m, k), k
    # file textures: mip 0 is the PNG (exact); the chain is rebuilt natively — the numpy builder in scenes.py may differ
    # from libm's powf by an ulp, which shows as at most 1 LSB in a handful of sRGB mip texels
    assert g.nfile_textures == sc.cubemap_index
    textures_identical = True
    for i in range(g.nfile_textures):
        data, offs, ws, hs, st = tex_arrays(d.textures[i])
        t = sc.textures[i]
        assert (d.textures[i].width, d.textures[i].height, d.textures[i].texture_type) == (t.width, t.height, t.texture_type)
        assert (d.textures[i].wrap_s, d.textures[i].wrap_t) == (t.wrap_s, t.wrap_t)
        assert np.array_equal(offs, t.mip_offsets) and np.array_equal(ws, t.mip_widths) and np.array_equal(hs, t.mip_heights) and np.array_equal(st, t.array_stride)
        n0 = t.width * t.height
        assert np.array_equal(data[:n0], t.data[:n0])
        diff = channel_diff(data, t.data)
        assert diff.max() <= 1 and np.count_nonzero(diff) <= max(4, diff.size // 500), (i, diff.max(), np.count_nonzero(diff))
        textures_identical &= bool(np.array_equal(data, t.data))
    # environment attached behind the file's textures, bit for bit
    for k in ("cubemap", "cubemap_specular", "brdf_lut"):
        a, b = tex_arrays(d.textures[getattr(d, k)]), tex_arrays(ref.textures[getattr(ref, k)])
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), k
    assert tuple(d.voxel_grid.dims[:]) == tuple(ref.voxel_grid.dims[:])
    nv = int(np.prod(d.voxel_grid.dims[:])) * 16
    assert np.array_equal(np.ctypeslib.as_array(d.voxel_grid.gi_sh4, (nv,)), np.ctypeslib.as_array(ref.voxel_grid.gi_sh4, (nv,)))
    # scene bounds (scene.rs:331-351: every vertex through its node transform)
    mn, mx = np.full(3, np.inf), np.full(3, -np.inf)
    for n in sc.nodes:
        M = n.transform.reshape(4, 4).T.astype(np.float64)
        for p in sc.meshes[n.mesh_index]:
            w = p.positions.astype(np.float64) @ M.T
            mn, mx = np.minimum(mn, w[:, :3].min(0)), np.maximum(mx, w[:, :3].max(0))
    assert np.allclose(g.bounds_min, mn, rtol=1e-5, atol=1e-5) and np.allclose(g.bounds_max, mx, rtol=1e-5, atol=1e-5)
    assert np.isclose(g.bounds_diagonal, np.linalg.norm(mx - mn), rtol=1e-5)
    assert g.total_triangles == sc.total_triangles
    # and the point of it all: the loaded scene renders to the same frame (oracle, CPU) when the inputs are identical
    if textures_identical:
        W, H = 192, 128
        cam = swr.RenderCamera.from_spec(spec, W, H)
        a, b = render_oracle(sc, cam, W, H), render_oracle(g, cam, W, H)
        for k in ("seq", "depth", "pixels"):
            assert np.array_equal(a[k], b[k]), k
    g.close()


def test_roundtrip_untextured_scene_renders_identically(tmp_path):
    sc, spec = scenes.scene_c1_sphere(segments=48, bands=32, **SMALL)
    scenes.export_gltf(sc, str(tmp_path / "sphere"))
    g = gltf.load_gltf(tmp_path / "sphere.gltf", environment=sc)
    W, H = 256, 160
    cam = swr.RenderCamera.from_spec(spec, W, H)
    a, b = render_oracle(sc, cam, W, H), render_oracle(g, cam, W, H)
    for k in ("seq", "depth", "pixels"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["seq"] != 0xFFFFFFFF).any()


# ---------------------------------------------------------------------------------------------------------------------
# (b) the loader's arithmetic against independent restatements
# ---------------------------------------------------------------------------------------------------------------------
def py_smooth_normals(pos, idx):
    """scene.rs:520-553 in scalar float32 Python (glam Vec3A: cross unfused, dot = (x*x + y*y) + z*z, normalize = v * (1/len))."""
    acc = [[F32(0)] * 3 for _ in range(len(pos))]
    for t in range(0, len(idx), 3):
        v0, v1, v2 = (pos[idx[t + k]] for k in range(3))
        e1 = [F32(v1[c] - v0[c]) for c in range(3)]
        e2 = [F32(v2[c] - v0[c]) for c in range(3)]
        fn = [F32(F32(e1[1] * e2[2]) - F32(e1[2] * e2[1])), F32(F32(e1[2] * e2[0]) - F32(e1[0] * e2[2])), F32(F32(e1[0] * e2[1]) - F32(e1[1] * e2[0]))]
        for k in range(3):
            a = acc[idx[t + k]]
            acc[idx[t + k]] = [F32(a[c] + fn[c]) for c in range(3)]
    out = np.zeros((len(pos), 4), F32)
    for i, a in enumerate(acc):
        l2 = F32(F32(F32(a[0] * a[0]) + F32(a[1] * a[1])) + F32(a[2] * a[2]))
        if l2 > 0:
            r = F32(F32(1.0) / F32(np.sqrt(l2)))
            out[i, :3] = [F32(a[c] * r) for c in range(3)]
        else:
            out[i, :3] = [0, 0, 1]
    return out


def py_tangents(pos, uv, nrm, idx):
    """scene.rs:555-646 in scalar float32 Python (glam Vec3: dot = x*x + y*y + z*z left to right)."""
    def sub(a, b): return [F32(a[c] - b[c]) for c in range(3)]
    def mul(a, s): return [F32(a[c] * s) for c in range(3)]
    def add(a, b): return [F32(a[c] + b[c]) for c in range(3)]
    def dot(a, b): return F32(F32(F32(a[0] * b[0]) + F32(a[1] * b[1])) + F32(a[2] * b[2]))
    def cross(a, b): return [F32(F32(a[1] * b[2]) - F32(a[2] * b[1])), F32(F32(a[2] * b[0]) - F32(a[0] * b[2])), F32(F32(a[0] * b[1]) - F32(a[1] * b[0]))]
    def normalize(a): return mul(a, F32(F32(1.0) / F32(np.sqrt(dot(a, a)))))
    n = len(pos)
    tan = [[F32(0)] * 3 for _ in range(n)]
    bit = [[F32(0)] * 3 for _ in range(n)]
    for t in range(0, len(idx), 3):
        i0, i1, i2 = int(idx[t]), int(idx[t + 1]), int(idx[t + 2])
        e1, e2 = sub(pos[i1][:3], pos[i0][:3]), sub(pos[i2][:3], pos[i0][:3])
        du1, dv1 = F32(uv[i1][0] - uv[i0][0]), F32(uv[i1][1] - uv[i0][1])
        du2, dv2 = F32(uv[i2][0] - uv[i0][0]), F32(uv[i2][1] - uv[i0][1])
        det = F32(F32(du1 * dv2) - F32(du2 * dv1))
        if abs(det) < F32(1e-6):
            continue
        tg = [F32(x / det) for x in sub(mul(e1, dv2), mul(e2, dv1))]
        bt = [F32(x / det) for x in sub(mul(e2, du1), mul(e1, du2))]
        for k in (i0, i1, i2):
            tan[k], bit[k] = add(tan[k], tg), add(bit[k], bt)
    out = np.zeros((n, 4), F32)
    for i in range(n):
        nn, t = [F32(x) for x in nrm[i][:3]], tan[i]
        if dot(t, t) < F32(1e-6):
            helper = [F32(0), F32(1), F32(0)] if abs(nn[0]) > F32(0.9) else [F32(1), F32(0), F32(0)]
            out[i, :3], out[i, 3] = normalize(cross(nn, helper)), 1.0
            continue
        tg = normalize(sub(t, mul(nn, dot(nn, t))))
        out[i, :3] = tg
        out[i, 3] = -1.0 if dot(cross(nn, tg), bit[i]) < 0 else 1.0
    return out


@pytest.mark.parametrize("mesh", ["torus", "sphere", "degenerate"])
def test_auto_normals_and_tangents_match_the_restatement(mesh):
    if mesh == "torus":
        p = scenes.torus(14, 10)
    elif mesh == "sphere":
        p = scenes.uv_sphere(12, 8)
    else:  # an isolated vertex, a zero-area triangle and constant uvs (fallback tangents, scene.rs:619-629)
        pos = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1], [5, 5, 5, 1], [0, 0, 1, 1]], F32)
        p = scenes.Primitive(pos, np.zeros((5, 4), F32), np.zeros((5, 4), F32), np.zeros((5, 2), F32), np.array([0, 1, 2, 0, 0, 1, 0, 4, 1], np.uint32), 0)
    with np.errstate(all="ignore"):
        want_n = py_smooth_normals(p.positions, p.indices)
        got_n = gltf.compute_smooth_normals(p.positions, p.indices)
        assert np.array_equal(got_n.view(np.uint32), want_n.view(np.uint32))
        want_t = py_tangents(p.positions, p.texcoords, want_n, p.indices)
        got_t = gltf.compute_tangents(p.positions, p.texcoords, got_n, p.indices)
    same = (got_t.view(np.uint32) == want_t.view(np.uint32)) | (np.isnan(got_t) & np.isnan(want_t))
    assert same.all()
    if mesh != "degenerate":
        assert np.allclose(np.linalg.norm(got_n[:, :3], axis=1), 1, atol=1e-5) and np.allclose(np.linalg.norm(got_t[:, :3], axis=1), 1, atol=1e-5)
        assert np.abs(np.einsum("ij,ij->i", got_n[:, :3], got_t[:, :3])).max() < 1e-4  # orthonormalised
        assert set(np.unique(got_t[:, 3])) <= {-1.0, 1.0}
    else:
        assert np.array_equal(got_n[3, :3], [0, 0, 1])  # vertex without faces (scene.rs:545-547)


def test_missing_normals_and_tangents_are_generated_on_load(tmp_path):
    p = scenes.torus(10, 8)
    pos3, uv, idx = np.ascontiguousarray(p.positions[:, :3]), p.texcoords, p.indices.astype(np.uint32)
    blob = pos3.tobytes() + uv.tobytes() + idx.tobytes()
    doc = {"asset": {"version": "2.0"}, "buffers": [{"uri": "mesh.bin", "byteLength": len(blob)}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": pos3.nbytes}, {"buffer": 0, "byteOffset": pos3.nbytes, "byteLength": uv.nbytes},
                           {"buffer": 0, "byteOffset": pos3.nbytes + uv.nbytes, "byteLength": idx.nbytes}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": len(pos3), "type": "VEC3"},
                         {"bufferView": 1, "componentType": 5126, "count": len(uv), "type": "VEC2"},
                         {"bufferView": 2, "componentType": 5125, "count": len(idx), "type": "SCALAR"}],
           "materials": [{"pbrMetallicRoughness": {"metallicFactor": 0.25}}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2}]}], "nodes": [{"mesh": 0}]}
    g = gltf.load_gltf(write_gltf(tmp_path, doc, blobs={"mesh.bin": blob}))
    pos, nrm, tan, uv2, idx2 = prim_arrays(g.desc(), 0)
    assert np.array_equal(nrm, gltf.compute_smooth_normals(p.positions, idx)) and np.array_equal(tan, gltf.compute_tangents(p.positions, uv, nrm, idx))
    assert np.array_equal(uv2, uv) and np.all(pos[:, 3] == 1.0)
    m = g.desc().materials[0]  # glTF defaults (scene.rs:710-787 through the gltf crate's defaults)
    assert (m.metallic_factor, m.roughness_factor, m.alpha_cutoff, m.occlusion_strength, m.flags) == (0.25, 1.0, 0.5, 1.0, 0)
    assert list(m.base_color_factor) == [1, 1, 1, 1] and m.base_color_texture == -1
    assert g.desc().primitives[0].material_index == 0  # material().index().unwrap_or(0)


def py_mip_chain(base, w, h, ttype):
    """Texture::generate_mipmaps (texture.rs:45-128) with math.pow-free float32 scalars where possible; returns all levels."""
    def unpack(p):
        return [F32(F32((p >> s) & 255) / F32(255.0)) for s in (24, 16, 8, 0)]

    def pack(c):
        out = 0
        for v, s in zip(c, (24, 16, 8, 0)):
            q = F32(v * F32(255.0))
            out |= (int(q) if q > 0 else 0) << s
        return out & 0xFFFFFFFF

    def s2l(x):
        return F32(x / F32(12.92)) if x <= F32(0.04045) else F32(np.power(F32(F32(x + F32(0.055)) / F32(1.055)), F32(2.4)))

    def l2s(x):
        return F32(x * F32(12.92)) if x <= F32(0.0031308) else F32(F32(np.power(x, F32(1.0 / 2.4)) * F32(1.055)) - F32(0.055))
    levels = [list(int(v) for v in base)]
    pw, ph = w, h
    nm = 1 + int(np.floor(np.log2(max(w, h))))
    for mip in range(1, nm):
        mw, mh = max(w >> mip, 1), max(h >> mip, 1)
        prev, cur = levels[-1], []
        for y in range(mh):
            for x in range(mw):
                x0, y0 = 2 * x, 2 * y
                x1, y1 = min(x0 + 1, pw - 1), min(y0 + 1, ph - 1)
                p = [unpack(prev[yy * pw + xx]) for yy, xx in ((y0, x0), (y0, x1), (y1, x0), (y1, x1))]
                if ttype == abi.TEX_SRGB:
                    avg = []
                    for c in range(4):
                        v = [s2l(q[c]) if c < 3 else q[c] for q in p]
                        m = F32(F32(F32(F32(v[0] + v[1]) + v[2]) + v[3]) / F32(4.0))
                        avg.append(l2s(m) if c < 3 else m)
                elif ttype == abi.TEX_METALLIC_ROUGHNESS:
                    r = F32(F32(F32(F32(F32(p[0][1] * p[0][1]) + F32(p[1][1] * p[1][1])) + F32(p[2][1] * p[2][1])) + F32(p[3][1] * p[3][1])) / F32(4.0))
                    m = F32(F32(F32(F32(p[0][2] + p[1][2]) + p[2][2]) + p[3][2]) / F32(4.0))
                    avg = [p[0][0], F32(np.sqrt(r)), m, p[0][3]]
                elif ttype == abi.TEX_NORMAL:
                    avg = []
                    for c in range(4):
                        v = [F32(F32(q[c] * F32(2.0)) - F32(1.0)) for q in p]
                        m = F32(F32(F32(F32(v[0] + v[1]) + v[2]) + v[3]) / F32(4.0))
                        avg.append(F32(F32(m + F32(1.0)) / F32(2.0)))
                else:
                    avg = [F32(F32(F32(F32(p[0][c] + p[1][c]) + p[2][c]) + p[3][c]) / F32(4.0)) for c in range(4)]
                cur.append(pack(avg))
        levels.append(cur)
        pw, ph = mw, mh
    return levels


@pytest.mark.parametrize("ttype", [abi.TEX_LINEAR, abi.TEX_NORMAL, abi.TEX_METALLIC_ROUGHNESS, abi.TEX_SRGB])
def test_mip_chain_matches_the_restatement(ttype):
    rng = np.random.default_rng(5 + ttype)
    w, h = 13, 6  # odd sizes: clamped neighbours (texture.rs:79-80), non-square tail down to 1x1
    base = rng.integers(0, 2 ** 32, w * h, dtype=np.uint64).astype(np.uint32)
    data, offs, ws, hs, st = gltf.build_mip_chain(base, w, h, ttype)
    levels = py_mip_chain(base, w, h, ttype)
    assert len(offs) == len(levels) == 4 and list(ws) == [13, 6, 3, 1] and list(hs) == [6, 3, 1, 1]
    assert list(st) == [0, 18, 3, 1]  # array_stride: 0 for mip 0 of a 2D texture, w*h after (texture.rs:64, :988)
    for lv, (o, ww, hh) in enumerate(zip(offs, ws, hs)):
        got, want = data[o:o + ww * hh], np.array(levels[lv], np.uint32)
        if ttype == abi.TEX_SRGB and lv > 0:  # numpy's powf and libm's may differ by an ulp -> at most 1 LSB after the truncation
            assert channel_diff(got, want).max() <= 1
        else:
            assert np.array_equal(got, want), (ttype, lv)
    assert o + ww * hh == len(data)


# ---------------------------------------------------------------------------------------------------------------------
# (c) container rules
# ---------------------------------------------------------------------------------------------------------------------
def test_glb_stride_sparse_normalised_and_hierarchy(tmp_path):
    # interleaved vertex buffer: position (12 B) + u16 normalised uv (4 B) = stride 16; u8 indices; sparse override of vertex 2
    pos = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0]], F32)
    uv16 = np.array([[0, 0], [65535, 0], [0, 65535], [32768, 65535]], np.uint16)
    inter = b"".join(pos[i].tobytes() + uv16[i].tobytes() for i in range(4))
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint8).tobytes() + b"\0\0"
    sp_idx = np.array([2], np.uint16).tobytes() + b"\0\0"
    sp_val = np.array([[0, 3, 1]], F32).tobytes()
    uv8 = np.array([[0, 0], [255, 0], [0, 255], [51, 255]], np.uint8).tobytes()
    chunks = [inter, idx, sp_idx, sp_val, uv8]
    offs = np.cumsum([0] + [len(c) for c in chunks])
    binary = b"".join(chunks)
    doc = {"asset": {"version": "2.0", "generator": "tést \"quoted\" \\ / €"}, "buffers": [{"byteLength": len(binary)}],
           "bufferViews": [{"buffer": 0, "byteOffset": int(offs[0]), "byteLength": len(inter), "byteStride": 16},
                           {"buffer": 0, "byteOffset": int(offs[1]), "byteLength": 6},
                           {"buffer": 0, "byteOffset": int(offs[2]), "byteLength": 2},
                           {"buffer": 0, "byteOffset": int(offs[3]), "byteLength": 12},
                           {"buffer": 0, "byteOffset": int(offs[4]), "byteLength": 8}],
           "accessors": [{"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3",
                          "sparse": {"count": 1, "indices": {"bufferView": 2, "componentType": 5123}, "values": {"bufferView": 3}}},
                         {"bufferView": 0, "byteOffset": 12, "componentType": 5123, "normalized": True, "count": 4, "type": "VEC2"},
                         {"bufferView": 1, "componentType": 5121, "count": 6, "type": "SCALAR"},
                         {"bufferView": 4, "componentType": 5121, "normalized": True, "count": 4, "type": "VEC2"}],
           "materials": [{"alphaMode": "MASK", "alphaCutoff": 2.5e-1, "emissiveFactor": [1e0, 0.5, 0.25],
                          "extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.75}}}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0},
                                      {"attributes": {"POSITION": 0, "TEXCOORD_0": 3}, "indices": 2}]}],
           "cameras": [{"type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1}},
                       {"type": "orthographic", "orthographic": {"xmag": 2, "ymag": 3, "znear": 0.5, "zfar": 9}}],
           "nodes": [{"children": [1, 2], "translation": [1, 2, 3], "scale": [2, 2, 2]},
                     {"mesh": 0, "rotation": [0, 0.7071067811865476, 0, 0.7071067811865476], "translation": [0, 0, -1], "children": [3]},
                     {"camera": 0, "translation": [0, 5, 0]},
                     {"mesh": 0, "matrix": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 4, 0, 0, 1], "camera": 1}],
           "scenes": [{"nodes": [0]}], "scene": 0, "extras": {"nested": [[[], {}], None, True, False, -1.5e-3]}}
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    binary_p = binary + b"\0" * (-len(binary) % 4)
    glb = b"glTF" + struct.pack("<II", 2, 12 + 8 + len(js) + 8 + len(binary_p)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(binary_p), 0x004E4942) + binary_p
    (tmp_path / "m.glb").write_bytes(glb)
    g = gltf.load_gltf(tmp_path / "m.glb")
    d = g.desc()
    assert (d.nprimitives, d.nmeshes, d.nnodes, d.nmaterials, d.ntextures) == (2, 1, 4, 1, 0)
    pos_l, nrm, tan, uv, idx_l = prim_arrays(d, 0)
    want_pos = pos.copy()
    want_pos[2] = [0, 3, 1]
    assert np.array_equal(pos_l[:, :3], want_pos) and list(idx_l) == [0, 1, 2, 2, 1, 3]
    assert np.array_equal(uv, (uv16.astype(F32) / F32(65535.0)))
    assert np.array_equal(prim_arrays(d, 1)[3], np.array([[0, 0], [255, 0], [0, 255], [51, 255]], F32) / F32(255.0))
    m = d.materials[0]
    assert m.flags == abi.MAT_TRANSLUCENT  # MASK without a base colour texture is not alpha tested (scene.rs:716-717)
    assert (m.alpha_cutoff, m.transmission, list(m.emissive_factor)) == (0.25, 0.75, [1.0, 0.5, 0.25])
    # hierarchy (scene.rs:356-380): final = parent * local, T * R * S
    def trs_m(t=(0, 0, 0), q=(0, 0, 0, 1), s=(1, 1, 1)):
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4)
        M[:3, :3] = R * np.array(s)
        M[:3, 3] = t
        return M
    L = [trs_m((1, 2, 3), s=(2, 2, 2)), trs_m((0, 0, -1), (0, 0.7071067811865476, 0, 0.7071067811865476)), trs_m((0, 5, 0)), trs_m((4, 0, 0))]
    want = [L[0], L[0] @ L[1], L[0] @ L[2], L[0] @ L[1] @ L[3]]
    for i in range(4):
        assert np.allclose(np.array(d.nodes[i].transform[:]).reshape(4, 4).T, want[i], atol=1e-6), i
    assert [d.nodes[i].mesh_index for i in range(4)] == [-1, 0, -1, 0]
    # node sphere centre is the LOCAL origin even for a child (scene.rs:384-386 takes the local transform)
    assert np.allclose(d.nodes[1].bounding_sphere_world[:3], [0, 0, -1]) and d.nodes[1].bounding_sphere_world[3] > 0
    assert d.nodes[0].bounding_sphere_world[3] == 0.0
    assert len(g.cameras) == 2 and g.cameras[0]["perspective"] and not g.cameras[1]["perspective"]
    assert (g.cameras[0]["yfov_or_xmag"], g.cameras[0]["aspect_or_ymag"], g.cameras[0]["zfar"]) == (0.5, 1.0, 100.0)  # unwrap_or defaults (scene.rs:793-794)
    assert np.allclose(g.cameras[0]["transform"].reshape(4, 4).T, want[2], atol=1e-6) and np.allclose(g.cameras[1]["transform"].reshape(4, 4).T, want[3], atol=1e-6)
    assert (g.cameras[1]["yfov_or_xmag"], g.cameras[1]["aspect_or_ymag"], g.cameras[1]["znear"], g.cameras[1]["zfar"]) == (2.0, 3.0, 0.5, 9.0)
    # bounds over both mesh nodes
    pts = np.concatenate([(np.c_[want_pos, np.ones(4)] @ want[i].T)[:, :3] for i in (1, 3)])
    assert np.allclose(g.bounds_min, pts.min(0), atol=1e-5) and np.allclose(g.bounds_max, pts.max(0), atol=1e-5)


def png_bytes(rgba, mode, interlace=False):
    from PIL import Image
    img = Image.fromarray(rgba, "RGBA").convert(mode) if mode != "RGBA" else Image.fromarray(rgba, "RGBA")
    buf = io.BytesIO()
    img.save(buf, "PNG", **({"interlace": 1} if interlace else {}))
    return buf.getvalue(), np.asarray(img.convert("RGBA"))


def raw_png(rgba, ctype, filters):
    """Minimal encoder with explicit per-row filter types (0..4) to exercise every reconstruction branch."""
    h, w, _ = rgba.shape
    rows = {6: rgba, 2: rgba[:, :, :3], 0: rgba[:, :, :1], 4: rgba[:, :, [0, 3]]}[ctype].astype(np.int32)
    bpp = rows.shape[2]
    rows = rows.reshape(h, -1)
    out = bytearray()
    prev = np.zeros(rows.shape[1], np.int32)
    for y in range(h):
        f = filters[y % len(filters)]
        cur = rows[y]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if f == 0:
            enc = cur
        elif f == 1:
            enc = cur - a
        elif f == 2:
            enc = cur - prev
        elif f == 3:
            enc = cur - ((a + prev) >> 1)
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            enc = cur - np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        out += bytes([f]) + (enc & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    comp = zlib.compress(bytes(out))
    half = len(comp) // 2  # two IDAT chunks
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + chunk(b"IDAT", comp[:half]) + chunk(b"IDAT", comp[half:]) +
            chunk(b"IEND", b""))


def test_png_decoder_all_colour_types_filters_and_interlace():
    rng = np.random.default_rng(77)
    rgba = rng.integers(0, 256, (11, 7, 4), dtype=np.uint8)
    for ctype in (6, 2, 0, 4):
        got = gltf.decode_png(raw_png(rgba, ctype, [0, 1, 2, 3, 4]))
        want = rgba.copy()
        if ctype == 2:
            want[:, :, 3] = 255
        elif ctype == 0:
            want = np.stack([rgba[:, :, 0]] * 3 + [np.full(rgba.shape[:2], 255, np.uint8)], -1)
        elif ctype == 4:
            want = np.stack([rgba[:, :, 0]] * 3 + [rgba[:, :, 3]], -1)
        assert np.array_equal(got, want), ctype
    for mode in ("RGBA", "RGB", "L", "LA", "P"):
        for inter in (False, True):
            data, want = png_bytes(rgba, mode, inter)
            assert np.array_equal(gltf.decode_png(data), want), (mode, inter)
    with pytest.raises(gltf.GltfError, match="neither a PNG nor a JPEG"):
        gltf.decode_png(b"GIF89a" + b"\0" * 32)
    bad = bytearray(raw_png(rgba, 6, [0]))
    bad[60] ^= 0xFF
    with pytest.raises(gltf.GltfError, match="inflate|corrupt"):
        gltf.decode_png(bytes(bad))


def test_textures_samplers_cache_and_registered_images(tmp_path):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    from PIL import Image
    Image.fromarray(img, "RGBA").save(tmp_path / "base color.png")
    doc = tri_doc(extra_root={
        "images": [{"uri": "base%20color.png"}, {"uri": "photo.jpg"}, {"bufferView": 0, "mimeType": "image/png"}],
        "samplers": [{"wrapS": 33071, "wrapT": 33648}, {}],
        "textures": [{"source": 0, "sampler": 0}, {"source": 0}, {"source": 1, "sampler": 1}, {"source": 2}, {"source": 0, "sampler": 0}],
        "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}}, "alphaMode": "MASK",
                       "normalTexture": {"index": 2}, "occlusionTexture": {"index": 3, "strength": 0.5}, "emissiveTexture": {"index": 4}}]})
    path = write_gltf(tmp_path, doc)
    with pytest.raises(gltf.GltfError, match="Could not load texture 'photo.jpg'"):
        gltf.load_gltf(path)
    jpg = rng.integers(0, 256, (4, 4, 4), dtype=np.uint8)
    gltf.register_image("photo.jpg", jpg)
    try:
        g = gltf.load_gltf(path)
    finally:
        gltf.register_image("photo.jpg", None)
    d, m = g.desc(), g.desc().materials[0]
    # slots: (uri, wrap) pairs in first-use order; an image without uri gives None (scene.rs:706); equal pairs share a slot
    assert (m.base_color_texture, m.metallic_roughness_texture, m.normal_texture, m.occlusion_texture, m.emissive_texture) == (0, 1, 2, -1, 0)
    assert m.flags == abi.MAT_ALPHA_TESTED and m.occlusion_strength == 0.5
    assert d.ntextures == 3 and g.texture_uri(0) == "base%20color.png" and g.texture_uri(2) == "photo.jpg"
    t0, t1, t2 = d.textures[0], d.textures[1], d.textures[2]
    assert (t0.wrap_s, t0.wrap_t) == (abi.WRAP_CLAMP_TO_EDGE, abi.WRAP_MIRRORED_REPEAT)
    assert (t1.wrap_s, t1.wrap_t) == (abi.WRAP_REPEAT, abi.WRAP_REPEAT) == (t2.wrap_s, t2.wrap_t)  # default sampler / empty sampler
    # TextureCache is keyed by URI: the first requested type (sRGB, from baseColorTexture) sticks for every later use
    assert t0.texture_type == abi.TEX_SRGB == t1.texture_type and t2.texture_type == abi.TEX_NORMAL
    assert C_ptr(t0.data) == C_ptr(t1.data)  # same texels shared, only the sampler differs
    data = tex_arrays(t0)[0]
    assert np.array_equal(data[:128], scenes.pack_rgba8(img).reshape(-1)) and t0.max_mip_level == 4 and (t0.width, t0.height) == (16, 8)
    assert np.array_equal(tex_arrays(t2)[0][:16], scenes.pack_rgba8(jpg).reshape(-1)) and t2.max_mip_level == 2


def C_ptr(p):
    import ctypes
    return ctypes.cast(p, ctypes.c_void_p).value


# ---------------------------------------------------------------------------------------------------------------------
# (d) errors
# ---------------------------------------------------------------------------------------------------------------------
def test_error_wording_follows_the_reference(tmp_path):
    doc = tri_doc()
    del doc["meshes"][0]["primitives"][0]["attributes"]["POSITION"]
    with pytest.raises(gltf.GltfError, match="Missing data: No positions in primitive"):  # scene.rs:446
        gltf.load_gltf(write_gltf(tmp_path, doc, "a.gltf"))
    doc = tri_doc()
    del doc["meshes"][0]["primitives"][0]["indices"]
    with pytest.raises(gltf.GltfError, match="Missing data: No indices in primitive"):  # scene.rs:452
        gltf.load_gltf(write_gltf(tmp_path, doc, "b.gltf"))
    with pytest.raises(gltf.GltfError, match="vertex index out of range"):
        gltf.load_gltf(write_gltf(tmp_path, tri_doc(indices=[0, 1, 3]), "c.gltf"))
    doc = tri_doc()
    doc["accessors"][0]["count"] = 40
    with pytest.raises(gltf.GltfError, match="outside its bufferView"):
        gltf.load_gltf(write_gltf(tmp_path, doc, "d.gltf"))
    doc = tri_doc()
    del doc["materials"]
    with pytest.raises(gltf.GltfError, match="no materials"):  # the reference would index materials[0] and panic (scene.rs:499)
        gltf.load_gltf(write_gltf(tmp_path, doc, "e.gltf"))
    (tmp_path / "f.gltf").write_text('{"asset": {"version": "2.0"}, "nodes": [}')
    with pytest.raises(gltf.GltfError, match="JSON"):
        gltf.load_gltf(tmp_path / "f.gltf")
    with pytest.raises(gltf.GltfError, match="cannot open"):
        gltf.load_gltf(tmp_path / "does_not_exist.gltf")
    doc = tri_doc()
    doc["nodes"] = [{"children": [1]}, {"children": [0], "mesh": 0}]
    g = gltf.load_gltf(write_gltf(tmp_path, doc, "g.gltf"))  # a cycle has no root: nothing is traversed, local transforms stay (scene.rs:300-316)
    assert g.desc().nnodes == 2
    doc["nodes"] = [{"children": [1, 1]}, {"mesh": 0}]
    with pytest.raises(gltf.GltfError, match="not a forest"):
        gltf.load_gltf(write_gltf(tmp_path, doc, "h.gltf"))


def test_hostile_sizes_are_rejected_not_trusted(tmp_path):
    """Offsets, counts and image headers come from the file: negative or absurd values must end in an error, never in an
    out-of-bounds read or a multi-gigabyte allocation (the loader was also byte- and field-fuzzed under ASan/UBSan)."""
    for k, (field, value) in enumerate([("byteOffset", -8), ("byteOffset", 2 ** 63), ("byteLength", -1), ("byteLength", 1e30), ("byteStride", -4), ("byteStride", 1e9)]):
        doc = tri_doc()
        doc["bufferViews"][0][field] = value
        with pytest.raises(gltf.GltfError, match="outside"):
            gltf.load_gltf(write_gltf(tmp_path, doc, f"v{k}.gltf"))
    for k, value in enumerate([-1, 2 ** 40, 1e300]):
        doc = tri_doc()
        doc["accessors"][0]["count"] = value
        with pytest.raises(gltf.GltfError, match="count out of range|outside"):
            gltf.load_gltf(write_gltf(tmp_path, doc, f"c{k}.gltf"))
    doc = tri_doc()
    doc["accessors"][0]["sparse"] = {"count": -3, "indices": {"bufferView": 1, "componentType": 5123}, "values": {"bufferView": 0}}
    with pytest.raises(gltf.GltfError, match="sparse count"):
        gltf.load_gltf(write_gltf(tmp_path, doc, "s.gltf"))
    doc = tri_doc()
    doc["nodes"] = [{"children": [i + 1]} for i in range(3000)] + [{"mesh": 0}]
    with pytest.raises(gltf.GltfError, match="deeper than"):
        gltf.load_gltf(write_gltf(tmp_path, doc, "deep.gltf"))
    doc["nodes"][5]["children"] = [-1]
    with pytest.raises(gltf.GltfError, match="child index"):
        gltf.load_gltf(write_gltf(tmp_path, doc, "neg.gltf"))
    rgba = np.zeros((4, 4, 4), np.uint8)
    png = bytearray(raw_png(rgba, 6, [0]))
    png[16:24] = struct.pack(">II", 30000, 30000)  # IHDR claims 3.6 GB of pixels for a few dozen compressed bytes (the CRC is not what stops it)
    with pytest.raises(gltf.GltfError, match="corrupt"):
        gltf.decode_png(bytes(png))
    png[16:24] = struct.pack(">II", 2 ** 31, 1)
    with pytest.raises(gltf.GltfError, match="larger than"):
        gltf.decode_png(bytes(png))


def raw_png_bits(samples, ctype, depth):
    """Encoder for any legal bit depth, filter 0: samples is (h, w, channels) of integers in file precision."""
    h, w, ch = samples.shape
    out = bytearray()
    for y in range(h):
        row = samples[y].reshape(-1).astype(np.uint32)
        if depth == 16:
            data = row.astype(">u2").tobytes()
        elif depth == 8:
            data = row.astype(np.uint8).tobytes()
        else:
            bits = "".join(format(int(v), f"0{depth}b") for v in row)
            bits += "0" * (-len(bits) % 8)
            data = int(bits, 2).to_bytes(len(bits) // 8, "big")
        out += b"\x00" + data

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(out))) + chunk(b"IEND", b"")


def test_png_decoder_other_bit_depths():
    """1/2/4-bit grey and palette, 16-bit grey / RGB / RGBA. 16 -> 8 bits as image::to_rgba8: (v + 128) / 257."""
    from PIL import Image
    rng = np.random.default_rng(21)
    w, h = 13, 5  # scanlines that do not end on a byte boundary
    for depth in (1, 2, 4):
        g = rng.integers(0, 1 << depth, (h, w, 1))
        got = gltf.decode_png(raw_png_bits(g, 0, depth))
        want8 = (g[..., 0] * (255 // ((1 << depth) - 1))).astype(np.uint8)
        assert np.array_equal(got, np.stack([want8] * 3 + [np.full((h, w), 255, np.uint8)], -1)), depth
    for depth, mode in ((16, None),):
        g = rng.integers(0, 65536, (h, w, 1))
        got = gltf.decode_png(raw_png_bits(g, 0, 16))
        assert np.array_equal(got[..., 0], ((g[..., 0] + 128) // 257).astype(np.uint8)) and (got[..., 3] == 255).all()
        rgb = rng.integers(0, 65536, (h, w, 3))
        got = gltf.decode_png(raw_png_bits(rgb, 2, 16))
        assert np.array_equal(got[..., :3], ((rgb + 128) // 257).astype(np.uint8))
        rgba = rng.integers(0, 65536, (h, w, 4))
        got = gltf.decode_png(raw_png_bits(rgba, 6, 16))
        assert np.array_equal(got, ((rgba + 128) // 257).astype(np.uint8))
        ga = rng.integers(0, 65536, (h, w, 2))
        got = gltf.decode_png(raw_png_bits(ga, 4, 16))
        assert np.array_equal(got[..., 0], ((ga[..., 0] + 128) // 257).astype(np.uint8)) and np.array_equal(got[..., 3], ((ga[..., 1] + 128) // 257).astype(np.uint8))
    # against PIL's own writer and reader: bilevel, small palettes (PIL packs them into 1/2/4 bits), 16-bit grey
    bil = Image.fromarray((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), "L").convert("1")
    buf = io.BytesIO()
    bil.save(buf, "PNG")
    assert np.array_equal(gltf.decode_png(buf.getvalue()), np.asarray(bil.convert("RGBA")))
    for ncol, bits in ((2, 1), (4, 2), (16, 4)):
        idx = rng.integers(0, ncol, (h, w)).astype(np.uint8)
        im = Image.fromarray(idx, "P")
        im.putpalette([int(v) for v in rng.integers(0, 256, ncol * 3)])
        buf = io.BytesIO()
        im.save(buf, "PNG", bits=bits)
        assert buf.getvalue()[24] == bits  # IHDR bit depth
        assert np.array_equal(gltf.decode_png(buf.getvalue()), np.asarray(im.convert("RGBA"))), bits
    g16 = rng.integers(0, 65536, (h, w)).astype(np.uint16)
    buf = io.BytesIO()
    Image.fromarray(g16).save(buf, "PNG")
    assert np.array_equal(gltf.decode_png(buf.getvalue())[..., 0], ((g16.astype(np.uint32) + 128) // 257).astype(np.uint8))
    with pytest.raises(gltf.GltfError, match="illegal bit depth"):
        gltf.decode_png(raw_png_bits(rng.integers(0, 4, (h, w, 3)), 2, 2))


def test_mutated_files_never_crash_the_loader(tmp_path):
    """Regression guard for the ASan/UBSan fuzzing done offline (40 k mutations): byte-level damage to a valid .gltf, .glb and
    .png must end in GltfError or in a successful load, never in a crash (a segfault would take the test process down)."""
    rng = np.random.default_rng(2024)
    sc, _ = scenes.scene_materials_test(**SMALL)
    scenes.export_gltf(sc, str(tmp_path / "seed"))
    seeds = {"gltf": (tmp_path / "seed.gltf").read_bytes(), "png": (tmp_path / "seed_tex0.png").read_bytes()}
    g = json.loads(seeds["gltf"])
    blob = (tmp_path / "seed.bin").read_bytes()
    for k in ("images", "textures", "samplers"):
        g.pop(k, None)
    for m in g["materials"]:
        for kk in [k for k in m.get("pbrMetallicRoughness", {}) if k.endswith("Texture")]:
            del m["pbrMetallicRoughness"][kk]
        for kk in [k for k in m if k.endswith("Texture")]:
            del m[kk]
    del g["buffers"][0]["uri"]
    js = json.dumps(g).encode()
    js += b" " * (-len(js) % 4)
    seeds["glb"] = b"glTF" + struct.pack("<II", 2, 20 + len(js) + 8 + len(blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
    outcomes = {"ok": 0, "rejected": 0}
    for it in range(240):
        kind = ("gltf", "glb", "png")[it % 3]
        data = bytearray(seeds[kind])
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(data)))
            op = int(rng.integers(0, 4))
            if op == 0:
                data[pos] ^= 1 << int(rng.integers(0, 8))
            elif op == 1:
                data[pos] = int(rng.integers(0, 256))
            elif op == 2:
                del data[pos:pos + int(rng.integers(1, 16))]
            else:
                data.insert(pos, b"0123456789-{}[],:\"e."[int(rng.integers(0, 20))])
        try:
            if kind == "png":
                gltf.decode_png(bytes(data))
            else:
                (tmp_path / f"case.{kind}").write_bytes(bytes(data))
                gltf.load_gltf(tmp_path / f"case.{kind}").close()
            outcomes["ok"] += 1
        except gltf.GltfError:
            outcomes["rejected"] += 1
    assert outcomes["ok"] + outcomes["rejected"] == 240 and outcomes["rejected"] > 50

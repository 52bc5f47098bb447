"""Procedural, seeded scenes in the layout the hot path consumes.

Everything here is load-time input construction (the reference's scene.rs /
texture.rs / gi.rs load path is out of scope); the arrays built
here are fed bit-identically to the oracle and to the CUDA path.

Layouts follow the reference: positions Vec4 (w=1, scene.rs:446), normals Vec3A
(16-byte stride), tangents Vec4, texcoords Vec2, u32 triangle lists; textures are
RGBA8 with R in bits 31..24 and all mips concatenated (texture.rs:28-42).
"""
import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import abi

F32 = np.float32


# ----------------------------------------------------------------------------------
# textures
# ----------------------------------------------------------------------------------
def pack_rgba8(rgba_u8):
    """(...,4) uint8 -> uint32 with R in the MSB (util.rs:98-100)."""
    c = rgba_u8.astype(np.uint32)
    return (c[..., 0] << 24) | (c[..., 1] << 16) | (c[..., 2] << 8) | c[..., 3]


def unpack_rgba8_f(data_u32):
    """uint32 -> (...,4) float32 in [0,1] (util.rs:83-89)."""
    d = data_u32.astype(np.uint32)
    out = np.stack([(d >> 24) & 0xFF, (d >> 16) & 0xFF, (d >> 8) & 0xFF, d & 0xFF], axis=-1).astype(F32)
    return out / F32(255.0)


def _srgb_to_linear(x):
    return np.where(x <= 0.04045, x / F32(12.92), np.power((x + F32(0.055)) / F32(1.055), F32(2.4))).astype(F32)


def _linear_to_srgb(x):
    return np.where(x <= 0.0031308, x * F32(12.92), np.power(np.maximum(x, 0), F32(1.0 / 2.4)) * F32(1.055) - F32(0.055)).astype(F32)


def _pack_vec4(v):
    """rgba8_pack_vec4 (util.rs:91-96): truncating (c*255) as u32 per channel."""
    q = np.clip(np.floor(v * F32(255.0)), 0, 255).astype(np.uint32)
    return (q[..., 0] << 24) | (q[..., 1] << 16) | (q[..., 2] << 8) | q[..., 3]


@dataclass
class Texture:
    """texture.rs:28-42 + sampler wrap modes."""
    data: np.ndarray
    width: int
    height: int
    texture_type: int
    mip_offsets: np.ndarray
    mip_widths: np.ndarray
    mip_heights: np.ndarray
    array_stride: np.ndarray
    wrap_s: int = abi.WRAP_REPEAT
    wrap_t: int = abi.WRAP_REPEAT

    @property
    def max_mip_level(self):
        return len(self.mip_offsets) - 1


def make_texture(base_u32, width, height, texture_type, wrap=abi.WRAP_REPEAT, slices=1, mips=True):
    """Mip chain as Texture::generate_mipmaps builds it (texture.rs:45-128), type-aware.

    base_u32: (slices, height, width) uint32.
    """
    base_u32 = np.asarray(base_u32, dtype=np.uint32).reshape(slices, height, width)
    levels = [base_u32]
    offs, ws, hs, strides = [0], [width], [height], [width * height if slices > 1 else 0]
    total = base_u32.size
    if mips:
        num_mips = 1 + int(math.floor(math.log2(max(width, height))))
        for mip in range(1, num_mips):
            prev = levels[-1]
            ph, pw = prev.shape[1], prev.shape[2]
            mw, mh = max(width >> mip, 1), max(height >> mip, 1)
            x0 = np.arange(mw) * 2
            y0 = np.arange(mh) * 2
            x1 = np.minimum(x0 + 1, pw - 1)
            y1 = np.minimum(y0 + 1, ph - 1)
            p = unpack_rgba8_f(prev)  # (s, ph, pw, 4)
            p00 = p[:, y0][:, :, x0]
            p10 = p[:, y0][:, :, x1]
            p01 = p[:, y1][:, :, x0]
            p11 = p[:, y1][:, :, x1]
            if texture_type == abi.TEX_SRGB:
                def lin(q):
                    r = q.copy()
                    r[..., :3] = _srgb_to_linear(q[..., :3])
                    return r
                avg = (lin(p00) + lin(p10) + lin(p01) + lin(p11)) / F32(4.0)
                out = avg.copy()
                out[..., :3] = _linear_to_srgb(avg[..., :3])
            elif texture_type == abi.TEX_METALLIC_ROUGHNESS:
                rough = (p00[..., 1] ** 2 + p10[..., 1] ** 2 + p01[..., 1] ** 2 + p11[..., 1] ** 2) / F32(4.0)
                metal = (p00[..., 2] + p10[..., 2] + p01[..., 2] + p11[..., 2]) / F32(4.0)
                out = np.stack([p00[..., 0], np.sqrt(rough), metal, p00[..., 3]], axis=-1)
            elif texture_type == abi.TEX_NORMAL:
                avg = ((p00 * 2 - 1) + (p10 * 2 - 1) + (p01 * 2 - 1) + (p11 * 2 - 1)) / F32(4.0)
                out = (avg + 1) / F32(2.0)
            else:
                out = (p00 + p10 + p01 + p11) / F32(4.0)
            lvl = _pack_vec4(out.astype(F32))
            offs.append(total)
            ws.append(mw)
            hs.append(mh)
            strides.append(mw * mh)
            total += lvl.size
            levels.append(lvl)
    data = np.concatenate([l.reshape(-1) for l in levels]).astype(np.uint32)
    return Texture(data, width, height, texture_type, np.array(offs, np.uint32), np.array(ws, np.uint32),
                   np.array(hs, np.uint32), np.array(strides, np.uint32), wrap, wrap)


def _value_noise(shape, cells, rng):
    """Bilinear value noise in [0,1] on an (h,w) grid with `cells` lattice cells per side (tileable)."""
    h, w = shape
    lat = rng.random((cells, cells)).astype(F32)
    ys = np.arange(h, dtype=F32) * (cells / h)
    xs = np.arange(w, dtype=F32) * (cells / w)
    y0 = np.floor(ys).astype(int)
    x0 = np.floor(xs).astype(int)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    y1 = (y0 + 1) % cells
    x1 = (x0 + 1) % cells
    y0 %= cells
    x0 %= cells
    a = lat[y0][:, x0]
    b = lat[y0][:, x1]
    c = lat[y1][:, x0]
    d = lat[y1][:, x1]
    return ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy).astype(F32)


def fbm(shape, rng, octaves=4, base_cells=4):
    out = np.zeros(shape, F32)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        out += F32(amp) * _value_noise(shape, base_cells << o, rng)
        tot += amp
        amp *= 0.5
    return out / F32(tot)


def checker_noise_texture(size, seed, texture_type=abi.TEX_SRGB, wrap=abi.WRAP_REPEAT, alpha_holes=False):
    """Seeded checker + noise RGBA8 texture with a full mip chain."""
    rng = np.random.default_rng(seed)
    n = fbm((size, size), rng, 4, 4)
    yy, xx = np.mgrid[0:size, 0:size]
    chk = (((xx * 16 // size) + (yy * 16 // size)) & 1).astype(F32)
    if texture_type == abi.TEX_NORMAL:
        gx = np.roll(n, -1, 1) - np.roll(n, 1, 1)
        gy = np.roll(n, -1, 0) - np.roll(n, 1, 0)
        nz = np.ones_like(n) * F32(2.0 / size * 8)
        nrm = np.stack([-gx, -gy, nz], -1)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        rgb = nrm * 0.5 + 0.5
    elif texture_type == abi.TEX_METALLIC_ROUGHNESS:
        rgb = np.stack([np.ones_like(n), 0.25 + 0.7 * n, 0.8 * chk], -1)
    elif texture_type == abi.TEX_LINEAR:
        rgb = np.stack([0.5 + 0.5 * n] * 3, -1)
    else:
        r = 0.25 + 0.6 * chk * n + 0.15 * n
        g = 0.30 + 0.5 * (1 - chk) * n + 0.2 * n
        b = 0.20 + 0.6 * n
        rgb = np.stack([r, g, b], -1)
    a = np.ones_like(n)
    if alpha_holes:
        a = (n > 0.45).astype(F32)
    rgba = np.clip(np.concatenate([rgb, a[..., None]], -1), 0, 1)
    u8 = np.floor(rgba * 255.0 + 0.5).astype(np.uint8)
    return make_texture(pack_rgba8(u8)[None], size, size, texture_type, wrap)


def _face_dirs(size):
    """Unit directions per cubemap texel, faces +X,-X,+Y,-Y,+Z,-Z (texture.rs:237-247)."""
    t = ((np.arange(size, dtype=F32) + 0.5) / size) * 2 - 1
    v, u = np.meshgrid(t, t, indexing="ij")
    one = np.ones_like(u)
    faces = [
        np.stack([one, -v, -u], -1), np.stack([-one, -v, u], -1), np.stack([u, one, v], -1),
        np.stack([u, -one, -v], -1), np.stack([u, -v, one], -1), np.stack([-u, -v, -one], -1),
    ]
    d = np.stack(faces, 0)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def sky_cubemap(size, seed):
    """Procedural sRGB sky: horizon gradient + sun + ground + noise. Returns (Texture CUBEMAP, float rgb faces)."""
    rng = np.random.default_rng(seed)
    d = _face_dirs(size)
    up = d[..., 1]
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.55 + 0.2 * (1 - up), 0.9 - 0.1 * (1 - up)], -1)
    ground = np.stack([0.32 + 0 * up, 0.28 + 0 * up, 0.22 + 0 * up], -1)
    t = np.clip(up * 4 + 0.5, 0, 1)[..., None]
    col = ground * (1 - t) + sky * t
    sun = np.array([-0.2, 1.0, 0.5], F32)
    sun /= np.linalg.norm(sun)
    s = np.clip((d @ sun), 0, 1) ** 64
    col = col + s[..., None] * np.array([1.0, 0.9, 0.7], F32)
    for f in range(6):
        col[f] += (fbm((size, size), rng, 3, 4)[..., None] - 0.5) * 0.08
    col = np.clip(col, 0, 1).astype(F32)
    rgba = np.concatenate([col, np.ones_like(col[..., :1])], -1)
    u8 = np.floor(rgba * 255 + 0.5).astype(np.uint8)
    tex = make_texture(pack_rgba8(u8), size, size, abi.TEX_CUBEMAP, abi.WRAP_CLAMP_TO_EDGE, slices=6)
    return tex, col


def specular_cubemap(col, size):
    """Stand-in for generate_prefiltered_specular_cubemap (texture.rs:330-420): every mip stays
    at full face resolution (:348-351); roughness progression approximated by a growing box blur
    per face. The GGX bake itself is load-time and out of scope."""
    from scipy.ndimage import uniform_filter
    lin = _srgb_to_linear(col)
    num_mips = 1 + int(math.floor(math.log2(size)))
    levels, offs, total = [], [], 0
    for mip in range(num_mips):
        r = 0 if mip == 0 else max(1, int(size * (mip / (num_mips - 1)) ** 2 * 0.25))
        img = lin if r == 0 else np.stack([uniform_filter(lin[f], size=(2 * r + 1, 2 * r + 1, 1), mode="nearest") for f in range(6)])
        rgba = np.concatenate([np.clip(img, 0, 1), np.ones_like(img[..., :1])], -1).astype(F32)
        lvl = _pack_vec4(rgba)
        offs.append(total)
        total += lvl.size
        levels.append(lvl.reshape(-1))
    data = np.concatenate(levels).astype(np.uint32)
    n = num_mips
    return Texture(data, size, size, abi.TEX_LINEAR, np.array(offs, np.uint32), np.full(n, size, np.uint32),
                   np.full(n, size, np.uint32), np.full(n, size * size, np.uint32), abi.WRAP_CLAMP_TO_EDGE, abi.WRAP_CLAMP_TO_EDGE)


def irradiance_sh4(col):
    """compute_irradiance_sh4 (texture.rs:289-328), nearest texel instead of bilinear."""
    size = col.shape[1]
    lin = _srgb_to_linear(col).astype(np.float64)
    d = _face_dirs(size).astype(np.float64)
    t = ((np.arange(size) + 0.5) / size) * 2 - 1
    v, u = np.meshgrid(t, t, indexing="ij")
    weight = ((2.0 / size) ** 2 / (1 + u * u + v * v) ** 1.5)[None, ..., None]
    basis = [0.282095 * np.ones_like(d[..., 0]), 0.488603 * d[..., 1], 0.488603 * d[..., 2], 0.488603 * d[..., 0]]
    sh = [np.sum(lin * (b[..., None] * weight), axis=(0, 1, 2)) for b in basis]
    sh[0] = sh[0] * math.pi * 0.282095
    for i in range(1, 4):
        sh[i] = sh[i] * (2 * math.pi / 3) * 0.488603
    return np.array(sh, F32)


def brdf_lut(size=128, samples=128):
    """generate_brdf_lut (texture.rs:199-235) vectorised in float64; Hammersley + GGX importance sampling. An INPUT builder, not
    a restatement: it leaves out the tangent frame of importance_sample_ggx (with n = +Z the reference's half vector is rotated by
    90 degrees about z), so single texels differ from the reference's table by a few LSB. The faithful version is
    host/swr_bake.hpp (tests/test_bakes.py)."""
    i = np.arange(samples, dtype=np.uint32)
    bits = i.copy()
    bits = (bits << 16) | (bits >> 16)
    bits = ((bits & 0x55555555) << 1) | ((bits & 0xAAAAAAAA) >> 1)
    bits = ((bits & 0x33333333) << 2) | ((bits & 0xCCCCCCCC) >> 2)
    bits = ((bits & 0x0F0F0F0F) << 4) | ((bits & 0xF0F0F0F0) >> 4)
    bits = ((bits & 0x00FF00FF) << 8) | ((bits & 0xFF00FF00) >> 8)
    xi_x = (i / samples).astype(np.float64)
    xi_y = bits.astype(np.float64) * 2.3283064e-10
    g = (np.arange(size) + 0.5) / size
    ndotv = np.maximum(g, 1e-4)[None, :, None]
    rough = np.maximum(g, 1e-4)[:, None, None]
    a = rough * rough
    phi = 2 * math.pi * xi_x[None, None, :]
    cos_t = np.sqrt((1 - xi_y) / (1 + (a * a - 1) * xi_y))
    sin_t = np.sqrt(np.maximum(1 - cos_t * cos_t, 0))
    hx, hz = np.cos(phi) * sin_t, cos_t
    vx, vz = np.sqrt(np.maximum(1 - ndotv * ndotv, 0)), ndotv
    vdoth = vx * hx + vz * hz
    lz = 2 * vdoth * hz - vz
    ndotl, ndoth, vdh = np.maximum(lz, 0), np.maximum(hz, 0), np.maximum(vdoth, 0)
    k = (a + 1) ** 2 * 0.125
    g_v = ndotv / (ndotv * (1 - k) + k)
    g_l = ndotl / (ndotl * (1 - k) + k)
    g_vis = np.maximum(g_v * g_l * vdh / (ndoth * np.maximum(ndotv, 1e-5) + 1e-30), 0)
    fc = (1 - vdh) ** 5
    m = ndotl > 0
    A = np.sum(np.where(m, (1 - fc) * g_vis, 0), -1) / samples
    B = np.sum(np.where(m, fc * g_vis, 0), -1) / samples
    rgba = np.stack([np.clip(A, 0, 1), np.clip(B, 0, 1), np.zeros_like(A), np.ones_like(A)], -1).astype(F32)
    return make_texture(_pack_vec4(rgba)[None], size, size, abi.TEX_LINEAR, abi.WRAP_CLAMP_TO_EDGE)


# ----------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------
@dataclass
class Primitive:
    positions: np.ndarray  # (n,4) f32
    normals: np.ndarray    # (n,4) f32 (Vec3A)
    tangents: np.ndarray   # (n,4) f32
    texcoords: np.ndarray  # (n,2) f32
    indices: np.ndarray    # (m,) u32
    material_index: int = 0

    def bounding_sphere(self):
        """compute_bounding_sphere (scene.rs:498-511)."""
        mn = self.positions[:, :3].min(0)
        mx = self.positions[:, :3].max(0)
        c = (mn + mx) * F32(0.5)
        r = F32(np.linalg.norm((mx - mn).astype(F32))) / F32(2.0)
        return np.array([c[0], c[1], c[2], r], F32)

    @property
    def ntris(self):
        return len(self.indices) // 3


def _finish_prim(pos3, nrm3, uv, idx, material=0, tangent3=None):
    n = len(pos3)
    pos = np.concatenate([pos3, np.ones((n, 1))], 1).astype(F32)
    nrm = np.concatenate([nrm3, np.zeros((n, 1))], 1).astype(F32)
    if tangent3 is None:
        ref = np.where(np.abs(nrm3[:, 1:2]) < 0.99, np.array([[0, 1, 0]], F32), np.array([[1, 0, 0]], F32))
        tangent3 = np.cross(ref, nrm3)
        tangent3 /= np.maximum(np.linalg.norm(tangent3, axis=1, keepdims=True), 1e-20)
    tan = np.concatenate([tangent3, np.ones((n, 1))], 1).astype(F32)
    return Primitive(np.ascontiguousarray(pos), np.ascontiguousarray(nrm), np.ascontiguousarray(tan),
                     np.ascontiguousarray(uv.astype(F32)), np.ascontiguousarray(idx.astype(np.uint32).reshape(-1)), material)


def uv_sphere(segments, bands, radius=1.0, material=0):
    """UV sphere with single-triangle pole caps: 2*S*(R-1) triangles. CCW from outside."""
    S, R = segments, bands
    theta = np.linspace(0, math.pi, R + 1)
    phi = np.linspace(0, 2 * math.pi, S + 1)
    th, ph = np.meshgrid(theta, phi, indexing="ij")
    n3 = np.stack([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    uv = np.stack([ph / (2 * math.pi), th / math.pi], -1).reshape(-1, 2)
    tan = np.stack([-np.sin(ph), np.zeros_like(ph), np.cos(ph)], -1).reshape(-1, 3)

    def vid(r, s):
        return r * (S + 1) + s
    tris = []
    r, s = np.meshgrid(np.arange(R), np.arange(S), indexing="ij")
    a, b, c, d = vid(r, s), vid(r + 1, s), vid(r + 1, s + 1), vid(r, s + 1)
    # quad (a: top-left, b: bottom-left, c: bottom-right, d: top-right); outward CCW = a, d, b / d, c, b
    t1 = np.stack([a, d, b], -1)
    t2 = np.stack([d, c, b], -1)
    top = r == 0      # a and d coincide at the pole: keep only t2
    bot = r == R - 1  # b and c coincide at the pole: keep only t1
    tris.append(t1[~top])
    tris.append(t2[~bot])
    idx = np.concatenate([t.reshape(-1, 3) for t in tris], 0)
    return _finish_prim(n3 * radius, n3, uv, idx, material, tan)


def icosphere(subdiv, material=0):
    """20 * 4^subdiv triangles, unit radius."""
    t = (1 + 5 ** 0.5) / 2
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        es = np.sort(e, 1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], 0)
        n = len(f)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], 0)
    uv = np.stack([np.arctan2(v[:, 2], v[:, 0]) / (2 * math.pi) + 0.5, np.arccos(np.clip(v[:, 1], -1, 1)) / math.pi], 1)
    return _finish_prim(v, v, uv, f, material)


def torus(nu, nv, R=1.0, r=0.4, material=0):
    """nu*nv*2 triangles."""
    u = np.arange(nu + 1) / nu * 2 * math.pi
    w = np.arange(nv + 1) / nv * 2 * math.pi
    uu, ww = np.meshgrid(u, w, indexing="ij")
    p = np.stack([(R + r * np.cos(ww)) * np.cos(uu), r * np.sin(ww), (R + r * np.cos(ww)) * np.sin(uu)], -1).reshape(-1, 3)
    n = np.stack([np.cos(ww) * np.cos(uu), np.sin(ww), np.cos(ww) * np.sin(uu)], -1).reshape(-1, 3)
    uv = np.stack([uu / (2 * math.pi) * 4, ww / (2 * math.pi)], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = i * (nv + 1) + j
    b = (i + 1) * (nv + 1) + j
    c = (i + 1) * (nv + 1) + j + 1
    d = i * (nv + 1) + j + 1
    idx = np.concatenate([np.stack([a, d, b], -1).reshape(-1, 3), np.stack([d, c, b], -1).reshape(-1, 3)], 0)
    return _finish_prim(p, n, uv, idx, material)


def box_grid(n, material=0):
    """Cube [-1,1]^3, each face an n x n grid: 6*n*n*2 triangles."""
    t = np.linspace(-1, 1, n + 1)
    a, b = np.meshgrid(t, t, indexing="ij")
    one = np.ones_like(a)
    faces = [
        (np.stack([one, a, -b], -1), [1, 0, 0]), (np.stack([-one, a, b], -1), [-1, 0, 0]),
        (np.stack([b, one, -a], -1), [0, 1, 0]), (np.stack([b, -one, a], -1), [0, -1, 0]),
        (np.stack([b, a, one], -1), [0, 0, 1]), (np.stack([-b, a, -one], -1), [0, 0, -1]),
    ]
    P, N, UV, I = [], [], [], []
    base = 0
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    for p, nrm in faces:
        P.append(p.reshape(-1, 3))
        N.append(np.tile(np.array(nrm, np.float64), ((n + 1) ** 2, 1)))
        UV.append(np.stack([(b + 1) / 2, (a + 1) / 2], -1).reshape(-1, 2))
        v00 = base + i * (n + 1) + j
        v01 = v00 + 1
        v10 = v00 + (n + 1)
        v11 = v10 + 1
        tri = np.concatenate([np.stack([v00, v01, v11], -1).reshape(-1, 3), np.stack([v00, v11, v10], -1).reshape(-1, 3)], 0)
        # orient CCW w.r.t. the outward normal
        pp = np.concatenate(P, 0) if False else None
        I.append(tri)
        base += (n + 1) ** 2
    P = np.concatenate(P, 0)
    N = np.concatenate(N, 0)
    I = np.concatenate(I, 0)
    e1 = P[I[:, 1]] - P[I[:, 0]]
    e2 = P[I[:, 2]] - P[I[:, 0]]
    flip = np.einsum("ij,ij->i", np.cross(e1, e2), N[I[:, 0]]) < 0
    I[flip] = I[flip][:, [0, 2, 1]]
    return _finish_prim(P, N, np.concatenate(UV, 0), I, material)


def height_field(nverts, seed, extent=10.0, height=1.2, uv_repeat=8.0, material=0, jitter=0.0, octaves=4):
    """(nverts-1)^2*2 triangle height-field in the xz plane, +y up."""
    rng = np.random.default_rng(seed)
    n = nverts
    h = fbm((n, n), rng, octaves, 4) * F32(height) if height != 0 else np.zeros((n, n), F32)
    t = np.linspace(-extent, extent, n)
    zz, xx = np.meshgrid(t, t, indexing="ij")
    if jitter > 0:
        step = 2 * extent / (n - 1)
        xx = xx + (rng.random((n, n)) - 0.5) * step * jitter
        zz = zz + (rng.random((n, n)) - 0.5) * step * jitter
    p = np.stack([xx, h, zz], -1).reshape(-1, 3)
    gy = np.gradient(h.astype(np.float64), t, axis=0)
    gx = np.gradient(h.astype(np.float64), t, axis=1)
    nrm = np.stack([-gx, np.ones_like(gx), -gy], -1).reshape(-1, 3)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = np.stack([(xx / (2 * extent) + 0.5) * uv_repeat, (zz / (2 * extent) + 0.5) * uv_repeat], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = i * n + j          # (z_i, x_j)
    b = (i + 1) * n + j    # z+1
    c = (i + 1) * n + j + 1
    d = i * n + j + 1      # x+1
    # +y facing CCW: a, b, d / d, b, c
    idx = np.concatenate([np.stack([a, b, d], -1).reshape(-1, 3), np.stack([d, b, c], -1).reshape(-1, 3)], 0)
    tan = np.tile(np.array([[1.0, 0, 0]]), (n * n, 1))
    return _finish_prim(p, nrm, uv, idx, material, tan)


@dataclass
class Material:
    base_color_factor: tuple = (0.8, 0.8, 0.8, 1.0)
    metallic_factor: float = 1.0
    roughness_factor: float = 1.0
    emissive_factor: tuple = (0.0, 0.0, 0.0)
    occlusion_strength: float = 1.0
    transmission: float = 0.0
    alpha_cutoff: float = 0.5
    flags: int = 0
    base_color_texture: int = -1
    metallic_roughness_texture: int = -1
    normal_texture: int = -1
    emissive_texture: int = -1
    occlusion_texture: int = -1
    transmission_texture: int = -1


@dataclass
class Node:
    transform: np.ndarray  # (16,) column-major f32
    mesh_index: int


@dataclass
class CameraSpec:
    position: tuple
    look_at: tuple
    fov: float
    far_plane: float


def trs(translate=(0, 0, 0), scale=1.0, rot_axis=(0, 1, 0), rot_angle=0.0):
    """Column-major local->world matrix with UNIFORM scale (renderer.rs:484 assumes it)."""
    ax = np.array(rot_axis, np.float64)
    ax /= np.linalg.norm(ax)
    c, s = math.cos(rot_angle), math.sin(rot_angle)
    x, y, z = ax
    R = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])
    M = np.eye(4)
    M[:3, :3] = R * scale
    M[:3, 3] = translate
    return np.ascontiguousarray(M.T.reshape(-1).astype(F32))  # column-major


class SceneData:
    """Owns the numpy arrays of a scene and exposes them as an abi.SceneDesc (mirror of scene.rs `Scene`)."""

    def __init__(self, meshes, nodes, materials, textures, cube_size=64, voxel_dim=16, seed=1, gi_variation=0.0,
                 light_direction=(-0.2, 1.0, 0.5), light_color=(5.0, 5.0, 4.75)):
        # meshes: list of list of Primitive
        self.meshes = meshes
        self.nodes = nodes
        self.materials = materials
        self.textures = list(textures)
        cube, col = sky_cubemap(cube_size, seed + 101)
        self.cubemap_index = len(self.textures)
        self.textures.append(cube)
        self.cubemap_specular_index = len(self.textures)
        self.textures.append(specular_cubemap(col, cube_size))
        self.brdf_lut_index = len(self.textures)
        self.textures.append(brdf_lut(128))
        self.irradiance_sh = irradiance_sh4(col)
        ld = np.array(light_direction, F32)
        self.light_direction = (ld / F32(np.linalg.norm(ld))).astype(F32)  # scene.rs:237
        self.light_color = np.array(light_color, F32)                       # (1,1,0.95)*5 scene.rs:238
        self._compute_bounds()
        self._build_voxel_grid(voxel_dim, seed, gi_variation)
        self._desc = None
        self._keep = []

    # -- bounds, spheres (scene.rs:321-353) -------------------------------------------
    def _compute_bounds(self):
        mn = np.full(3, np.inf, F32)
        mx = np.full(3, -np.inf, F32)
        self.node_spheres = []
        for nd in self.nodes:
            M = nd.transform.reshape(4, 4).T.astype(F32)
            radius_acc = F32(0.0)
            # Node::from_gltf (scene.rs:384-386): the sphere starts at the node's LOCAL origin with radius 0 (all nodes of the
            # procedural scenes are roots, so local == world); grow_by_sphere (scene.rs:44-49) then only grows the radius
            centre = (M @ np.array([0, 0, 0, 1], F32))[:3].astype(F32)
            if nd.mesh_index >= 0:
                for prim in self.meshes[nd.mesh_index]:
                    # bounds of the transformed bbox corners are enough for inputs (exact per-vertex for small meshes)
                    p = prim.positions
                    lo, hi = p[:, :3].min(0), p[:, :3].max(0)
                    corners = np.array([[x, y, z, 1] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], F32)
                    w = corners @ M.T
                    mn = np.minimum(mn, w[:, :3].min(0))
                    mx = np.maximum(mx, w[:, :3].max(0))
                    bs = prim.bounding_sphere()
                    # Mat4 * &BoundingSphere (scene.rs:53-63)
                    max_scale = (np.linalg.norm(M[:, 0]) + np.linalg.norm(M[:, 1]) + np.linalg.norm(M[:, 2])) / F32(3.0)
                    c = (M @ np.array([bs[0], bs[1], bs[2], 1], F32))[:3]
                    r = bs[3] * max_scale
                    dist = F32(np.linalg.norm((centre - c).astype(F32)))
                    if dist + r > radius_acc:
                        radius_acc = F32(dist + r)
            self.node_spheres.append(np.array([centre[0], centre[1], centre[2], radius_acc], F32))
        self.bounds_min, self.bounds_max = mn.astype(F32), mx.astype(F32)
        self.bounds_center = ((mn + mx) * F32(0.5)).astype(F32)
        self.bounds_diagonal = F32(np.linalg.norm((mx - mn).astype(F32)))

    def default_camera(self):
        """main.rs:210-224."""
        c, d = self.bounds_center, float(self.bounds_diagonal)
        return CameraSpec((float(c[0]), float(c[1]), float(c[2]) + d), tuple(float(x) for x in c), math.pi / 4, d * 2.0)

    # -- voxel grid (main.rs:228-235, gi.rs:123-149) ------------------------------------
    def _build_voxel_grid(self, dim, seed, gi_variation):
        rng = np.random.default_rng(seed + 7)
        self.voxel_dims = (dim, dim, dim)
        z, y, x = np.meshgrid(np.arange(dim), np.arange(dim), np.arange(dim), indexing="ij")
        # stand-in for compute_sun_visibility (gi.rs:267-314, load-time ray casting): smooth seeded field in [0.2, 1]
        ph = rng.random(3) * 6.28
        li = 0.6 + 0.4 * np.sin(x / dim * 5.1 + ph[0]) * np.cos(y / dim * 4.3 + ph[1]) * np.sin(z / dim * 3.7 + ph[2])
        g = np.zeros((dim, dim, dim, 4, 4), F32)
        for i in range(4):
            g[..., i, :3] = self.irradiance_sh[i] * F32(0.25)  # GI_FALLBACK_SCENE_SH_SCALE main.rs:54
        if gi_variation > 0:
            g[..., :, :3] *= (1 + gi_variation * (rng.random((dim, dim, dim, 1, 1)).astype(F32) - 0.5))
        g[..., 0, 3] = li.astype(F32)
        g[..., 1, 3] = 1.0
        self.gi_sh4 = np.ascontiguousarray(g.reshape(-1))

    # -- totals -----------------------------------------------------------------------------
    @property
    def total_triangles(self):
        return sum(sum(p.ntris for p in self.meshes[n.mesh_index]) for n in self.nodes if n.mesh_index >= 0)

    # -- abi -----------------------------------------------------------------------------------
    def desc(self):
        if self._desc is not None:
            return self._desc
        keep = self._keep

        def fp(a):
            a = np.ascontiguousarray(a, dtype=F32)
            keep.append(a)
            return a.ctypes.data_as(abi.f32p)

        def up(a):
            a = np.ascontiguousarray(a, dtype=np.uint32)
            keep.append(a)
            return a.ctypes.data_as(abi.u32p)

        prims, meshes = [], []
        for m in self.meshes:
            meshes.append(abi.MeshDesc(len(prims), len(m)))
            for p in m:
                d = abi.PrimitiveDesc()
                d.positions, d.normals, d.tangents, d.texcoords = fp(p.positions), fp(p.normals), fp(p.tangents), fp(p.texcoords)
                d.indices = up(p.indices)
                d.nverts, d.nindices, d.material_index = len(p.positions), len(p.indices), p.material_index
                d.bounding_sphere = (C.c_float * 4)(*p.bounding_sphere())
                prims.append(d)
        nodes = []
        for nd, sph in zip(self.nodes, self.node_spheres):
            d = abi.NodeDesc()
            d.transform = (C.c_float * 16)(*nd.transform)
            d.mesh_index = nd.mesh_index
            d.bounding_sphere_world = (C.c_float * 4)(*sph)
            nodes.append(d)
        mats = []
        for m in self.materials:
            d = abi.MaterialDesc()
            d.base_color_factor = (C.c_float * 4)(*m.base_color_factor)
            d.metallic_factor, d.roughness_factor = m.metallic_factor, m.roughness_factor
            d.emissive_factor = (C.c_float * 3)(*m.emissive_factor)
            d.occlusion_strength, d.transmission, d.alpha_cutoff, d.flags = m.occlusion_strength, m.transmission, m.alpha_cutoff, m.flags
            for k in ("base_color_texture", "metallic_roughness_texture", "normal_texture", "emissive_texture",
                      "occlusion_texture", "transmission_texture"):
                setattr(d, k, getattr(m, k))
            mats.append(d)
        texs = []
        for t in self.textures:
            d = abi.TextureDesc()
            d.data, d.ntexels = up(t.data), len(t.data)
            d.width, d.height, d.texture_type, d.max_mip_level = t.width, t.height, t.texture_type, t.max_mip_level
            d.mip_offsets, d.mip_widths, d.mip_heights, d.array_stride = up(t.mip_offsets), up(t.mip_widths), up(t.mip_heights), up(t.array_stride)
            d.wrap_s, d.wrap_t = t.wrap_s, t.wrap_t
            texs.append(d)
        sd = abi.SceneDesc()
        arr = lambda T, xs: (T * max(len(xs), 1))(*xs)
        self._prims, self._meshes, self._nodes, self._mats, self._texs = (arr(abi.PrimitiveDesc, prims), arr(abi.MeshDesc, meshes),
                                                                          arr(abi.NodeDesc, nodes), arr(abi.MaterialDesc, mats), arr(abi.TextureDesc, texs))
        sd.primitives, sd.nprimitives = self._prims, len(prims)
        sd.meshes, sd.nmeshes = self._meshes, len(meshes)
        sd.nodes, sd.nnodes = self._nodes, len(nodes)
        sd.materials, sd.nmaterials = self._mats, len(mats)
        sd.textures, sd.ntextures = self._texs, len(texs)
        sd.voxel_grid.dims = (C.c_uint32 * 3)(*self.voxel_dims)
        sd.voxel_grid.world_min = (C.c_float * 3)(*self.bounds_min)
        sd.voxel_grid.world_max = (C.c_float * 3)(*self.bounds_max)
        sd.voxel_grid.gi_sh4 = fp(self.gi_sh4)
        sd.cubemap, sd.cubemap_specular, sd.brdf_lut = self.cubemap_index, self.cubemap_specular_index, self.brdf_lut_index
        sd.light_direction = (C.c_float * 3)(*self.light_direction)
        sd.light_color = (C.c_float * 3)(*self.light_color)
        self._desc = sd
        return sd


IDENT = trs()


# ----------------------------------------------------------------------------------
# the BASELINE.json configurations (and scaled-down versions for tests)
# ----------------------------------------------------------------------------------
def scene_c1_sphere(segments=224, bands=224, **kw):
    """C1: UV sphere, untextured, one default material, identity node; default camera."""
    sc = SceneData([[uv_sphere(segments, bands)]], [Node(IDENT, 0)], [Material()], [], **kw)
    return sc, sc.default_camera()


def scene_c2_terrain(nverts=708, tex_size=2048, seed=0x5EED0002, **kw):
    """C2: textured height-field, low-oblique camera so several mips are hit."""
    tex = checker_noise_texture(tex_size, seed)
    mat = Material(base_color_factor=(1, 1, 1, 1), metallic_factor=0.1, roughness_factor=0.7, base_color_texture=0)
    sc = SceneData([[height_field(nverts, seed, extent=10.0, height=1.5, uv_repeat=6.0)]], [Node(IDENT, 0)], [mat], [tex], seed=seed & 0xFFFF, **kw)
    cam = CameraSpec((0.0, 2.6, 11.5), (0.0, 0.3, 0.0), math.pi / 4, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


def scene_c3_instanced(target_tris=10_000_000, seed=0x5EED0003, ico_subdiv=5, torus_n=64, box_n=10, aspect=16 / 9, **kw):
    """C3: three base meshes instanced by a seeded scatter with uniform scales in [0.05, 4];
    camera inside the cloud so a large share of instances straddle the frustum."""
    rng = np.random.default_rng(seed)
    meshes = [[icosphere(ico_subdiv, 0)], [torus(torus_n, torus_n, material=1)], [box_grid(box_n, 2)]]
    mats = [Material((0.85, 0.35, 0.25, 1), 0.0, 0.55), Material((0.9, 0.8, 0.4, 1), 1.0, 0.35), Material((0.3, 0.5, 0.85, 1), 0.2, 0.8)]
    tri_counts = [m[0].ntris for m in meshes]
    nodes, total = [], 0
    fov = math.pi / 4
    ty = math.tan(fov / 2)
    tx = ty * aspect
    dmax = 120.0
    while total < target_tris:
        mi = int(rng.integers(0, 3))
        scale = float(rng.uniform(0.05, 4.0))
        u = rng.random()
        if u < 0.06:  # around / behind the camera: straddles near plane or is culled
            pos = rng.uniform(-6, 6, 3)
        else:
            d = dmax * rng.random() ** (1 / 3)
            pos = np.array([rng.uniform(-1.2, 1.2) * tx * d, rng.uniform(-1.2, 1.2) * ty * d, -d])
        axis = rng.normal(size=3)
        nodes.append(Node(trs(pos, scale, axis, float(rng.uniform(0, 6.28))), mi))
        total += tri_counts[mi]
    sc = SceneData(meshes, nodes, mats, [], seed=seed & 0xFFFF, **kw)
    cam = CameraSpec((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), fov, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


def scene_c4_micro(nverts=5001, seed=0x5EED0004, aspect=16 / 9, **kw):
    """C4: one dense jittered grid filling the view (sub-pixel triangles, binning-bound)."""
    prim = height_field(nverts, seed, extent=1.0, height=0.02, uv_repeat=1.0, jitter=0.3, octaves=3)
    # grid lies in xz; rotate it to face the camera (+y -> +z) and stretch to the view's aspect via distance
    M = trs((0, 0, 0), 1.0, (1, 0, 0), math.pi / 2)
    sc = SceneData([[prim]], [Node(M, 0)], [Material((0.7, 0.7, 0.75, 1), 0.0, 0.6)], [], seed=seed & 0xFFFF, **kw)
    fov = math.pi / 4
    dist = 1.0 / (math.tan(fov / 2) * aspect) * 0.98  # horizontal extent slightly overfills the view
    cam = CameraSpec((0.0, 0.0, dist), (0.0, 0.0, 0.0), fov, 10.0)
    return sc, cam


def scene_c5_shards(nshards=8, nverts=3537, seed=0x5EED0005, aspect=16 / 9, **kw):
    """C5: nshards grids at different depths/offsets, one draw each (sharded by primitive, depth-composited)."""
    rng = np.random.default_rng(seed)
    meshes, nodes = [], []
    for s in range(nshards):
        meshes.append([height_field(nverts, seed + s, extent=1.0, height=0.15, uv_repeat=1.0, jitter=0.2, octaves=3, material=s % 3)])
        M = trs((float(rng.uniform(-0.15, 0.15)), float(rng.uniform(-0.1, 0.1)), -0.12 * s), 1.0, (1, 0, 0), math.pi / 2)
        nodes.append(Node(M, s))
    mats = [Material((0.8, 0.4, 0.3, 1), 0.0, 0.6), Material((0.4, 0.8, 0.3, 1), 0.3, 0.5), Material((0.3, 0.4, 0.8, 1), 0.6, 0.4)]
    sc = SceneData(meshes, nodes, mats, [], seed=seed & 0xFFFF, **kw)
    fov = math.pi / 4
    dist = 1.0 / (math.tan(fov / 2) * aspect) * 1.05
    cam = CameraSpec((0.0, 0.0, dist), (0.0, 0.0, -0.4), fov, 10.0)
    return sc, cam


def scene_materials_test(seed=11, tex_size=64, **kw):
    """Small scene exercising every shader branch: all five texture slots, three wrap modes, emissive."""
    texs = [
        checker_noise_texture(tex_size, seed, abi.TEX_SRGB, abi.WRAP_REPEAT),
        checker_noise_texture(tex_size, seed + 1, abi.TEX_METALLIC_ROUGHNESS, abi.WRAP_MIRRORED_REPEAT),
        checker_noise_texture(tex_size, seed + 2, abi.TEX_NORMAL, abi.WRAP_REPEAT),
        checker_noise_texture(tex_size // 2, seed + 3, abi.TEX_SRGB, abi.WRAP_CLAMP_TO_EDGE),
        checker_noise_texture(tex_size, seed + 4, abi.TEX_LINEAR, abi.WRAP_REPEAT),
    ]
    mats = [
        Material((1, 1, 1, 1), 1.0, 1.0, (0.6, 0.5, 0.2), 0.8, base_color_texture=0, metallic_roughness_texture=1,
                 normal_texture=2, emissive_texture=3, occlusion_texture=4),
        Material((0.9, 0.2, 0.2, 1), 0.0, 0.3),
        Material((0.2, 0.9, 0.3, 1), 1.0, 0.1, (0.05, 0.0, 0.1)),
    ]
    meshes = [[height_field(24, seed, extent=4.0, height=0.8, uv_repeat=3.0, material=0)], [uv_sphere(24, 16, 1.0, 1)], [torus(24, 12, material=2)]]
    nodes = [Node(IDENT, 0), Node(trs((-1.5, 1.4, 0.5), 0.9), 1), Node(trs((1.6, 1.3, -0.4), 1.1, (1, 0.3, 0.2), 0.8), 2),
             Node(trs((0.0, 1.0, 3.2), 0.7, (0, 1, 0), 0.4), 2)]
    sc = SceneData(meshes, nodes, mats, texs, seed=seed, gi_variation=0.5, **kw)
    cam = CameraSpec((0.5, 3.0, 6.5), (0.0, 0.8, 0.0), math.pi / 4, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


def scene_translucent_test(seed=31, **kw):
    """Opaque terrain + objects with KHR_materials_transmission objects in front of, between and intersecting them:
    exercises the per-tile back-to-front forward pass (tilerasterizer.rs:92-101, shader.rs:66-76, 265-277)."""
    texs = [checker_noise_texture(64, seed, abi.TEX_SRGB, abi.WRAP_REPEAT), checker_noise_texture(32, seed + 1, abi.TEX_LINEAR, abi.WRAP_REPEAT)]
    mats = [
        Material((1, 1, 1, 1), 0.0, 0.8, base_color_texture=0),
        Material((0.9, 0.3, 0.2, 1), 0.2, 0.4),
        Material((0.6, 0.9, 1.0, 1), 0.0, 0.15, transmission=0.8, flags=abi.MAT_TRANSLUCENT),
        Material((1.0, 0.8, 0.5, 1), 0.0, 0.3, transmission=0.6, flags=abi.MAT_TRANSLUCENT, transmission_texture=1, base_color_texture=0),
    ]
    meshes = [[height_field(24, seed, extent=4.0, height=0.7, uv_repeat=2.0, material=0)], [uv_sphere(20, 14, 1.0, 1)],
              [uv_sphere(24, 16, 1.0, 2)], [torus(28, 14, material=3), box_grid(3, 1)]]
    nodes = [Node(IDENT, 0), Node(trs((-0.4, 1.0, -0.5), 0.8), 1), Node(trs((0.4, 1.2, 1.4), 1.0), 2),
             Node(trs((-1.3, 1.3, 1.8), 0.9, (1, 0.2, 0.1), 0.9), 3), Node(trs((1.4, 1.0, 0.2), 0.7, (0, 1, 0), 0.5), 2),
             Node(trs((0.2, 1.6, 5.2), 1.6), 2)]  # a glass sphere crossing the near plane
    sc = SceneData(meshes, nodes, mats, texs, seed=seed, gi_variation=0.3, **kw)
    cam = CameraSpec((0.3, 2.6, 6.2), (0.0, 0.9, 0.0), math.pi / 4, float(sc.bounds_diagonal) * 2.0)
    return sc, cam


# ----------------------------------------------------------------------------------
# glTF 2.0 export: lets anyone with a Rust toolchain load these procedural scenes into the real swraster-viewer
# (scene.rs:145-354 reads POSITION / NORMAL / TANGENT / TEXCOORD_0, u32 indices, pbrMetallicRoughness factors, node
# matrices). Textures are referenced by URI and written as PNG when Pillow is available.
# ----------------------------------------------------------------------------------
def export_gltf(scene, path):
    """Write `scene` as <path>.gltf + <path>.bin (+ PNGs). Returns the glTF dict."""
    import json
    import os
    blobs, views, accessors = [], [], []
    offset = 0

    def add(arr, target, comp, typ, minmax=False):
        nonlocal offset
        raw = np.ascontiguousarray(arr).tobytes()
        pad = (-len(raw)) % 4
        views.append({"buffer": 0, "byteOffset": offset, "byteLength": len(raw), "target": target})
        acc = {"bufferView": len(views) - 1, "componentType": comp, "count": int(arr.shape[0]), "type": typ}
        if minmax:
            acc["min"] = [float(x) for x in arr.min(0)]
            acc["max"] = [float(x) for x in arr.max(0)]
        accessors.append(acc)
        blobs.append(raw + b"\0" * pad)
        offset += len(raw) + pad
        return len(accessors) - 1

    FLOAT, UINT, ARRAY, ELEMENT = 5126, 5125, 34962, 34963
    base = os.path.splitext(path)[0]
    images, textures, samplers = [], [], []
    wrap_code = {abi.WRAP_REPEAT: 10497, abi.WRAP_MIRRORED_REPEAT: 33648, abi.WRAP_CLAMP_TO_EDGE: 33071}
    user_textures = scene.textures[:scene.cubemap_index]  # cubemap / prefiltered / LUT are built-ins of the viewer
    for i, t in enumerate(user_textures):
        uri = f"{os.path.basename(base)}_tex{i}.png"
        try:
            from PIL import Image
            px = t.data[: t.width * t.height].reshape(t.height, t.width)
            rgba = np.stack([(px >> 24) & 255, (px >> 16) & 255, (px >> 8) & 255, px & 255], -1).astype(np.uint8)
            Image.fromarray(rgba, "RGBA").save(os.path.join(os.path.dirname(path) or ".", uri))
        except ImportError:
            pass
        images.append({"uri": uri})
        samplers.append({"wrapS": wrap_code[t.wrap_s], "wrapT": wrap_code[t.wrap_t]})
        textures.append({"source": i, "sampler": i})
    materials = []
    for m in scene.materials:
        pbr = {"baseColorFactor": [float(x) for x in m.base_color_factor], "metallicFactor": float(m.metallic_factor),
               "roughnessFactor": float(m.roughness_factor)}
        g = {"pbrMetallicRoughness": pbr, "emissiveFactor": [float(x) for x in m.emissive_factor]}
        if m.base_color_texture >= 0:
            pbr["baseColorTexture"] = {"index": m.base_color_texture}
        if m.metallic_roughness_texture >= 0:
            pbr["metallicRoughnessTexture"] = {"index": m.metallic_roughness_texture}
        if m.normal_texture >= 0:
            g["normalTexture"] = {"index": m.normal_texture}
        if m.emissive_texture >= 0:
            g["emissiveTexture"] = {"index": m.emissive_texture}
        if m.occlusion_texture >= 0:
            g["occlusionTexture"] = {"index": m.occlusion_texture, "strength": float(m.occlusion_strength)}
        if m.flags & abi.MAT_ALPHA_TESTED:
            g["alphaMode"], g["alphaCutoff"] = "MASK", float(m.alpha_cutoff)
        if m.flags & abi.MAT_TRANSLUCENT:
            ext = {"transmissionFactor": float(m.transmission)}
            if m.transmission_texture >= 0:
                ext["transmissionTexture"] = {"index": m.transmission_texture}
            g["extensions"] = {"KHR_materials_transmission": ext}
        materials.append(g)
    meshes = []
    for prims in scene.meshes:
        gp = []
        for p in prims:
            attrs = {"POSITION": add(p.positions[:, :3].astype(F32), ARRAY, FLOAT, "VEC3", True),
                     "NORMAL": add(p.normals[:, :3].astype(F32), ARRAY, FLOAT, "VEC3"),
                     "TANGENT": add(p.tangents.astype(F32), ARRAY, FLOAT, "VEC4"),
                     "TEXCOORD_0": add(p.texcoords.astype(F32), ARRAY, FLOAT, "VEC2")}
            gp.append({"attributes": attrs, "indices": add(p.indices.astype(np.uint32), ELEMENT, UINT, "SCALAR"),
                       "material": int(p.material_index), "mode": 4})
        meshes.append({"primitives": gp})
    nodes = [{"mesh": int(n.mesh_index), "matrix": [float(x) for x in n.transform]} for n in scene.nodes]
    gltf = {"asset": {"version": "2.0", "generator": "swraster-viewer_b200 scenes.py"}, "scene": 0,
            "scenes": [{"nodes": list(range(len(nodes)))}], "nodes": nodes, "meshes": meshes, "materials": materials,
            "accessors": accessors, "bufferViews": views, "buffers": [{"uri": os.path.basename(base) + ".bin", "byteLength": offset}]}
    if textures:
        gltf.update(textures=textures, images=images, samplers=samplers)
    if any(m.flags & abi.MAT_TRANSLUCENT for m in scene.materials):
        gltf["extensionsUsed"] = ["KHR_materials_transmission"]
    with open(base + ".bin", "wb") as f:
        f.write(b"".join(blobs))
    with open(base + ".gltf", "w") as f:
        json.dump(gltf, f)
    return gltf

"""The CHECKER of the device environment bakes (oracle/oracle_bake.cpp: CPU restatement of texture.rs:135-420) against
analytic properties and independent restatements written here in float64 numpy / scalar Python. CPU only: this is what pins
the oracle that tests/test_gpu_bakes.py then holds the CUDA bakes (csrc/swr_bake.cuh) to.
Tolerances are stated where float32-vs-float64 evaluation or libm differences can move a value across an 8-bit boundary."""
import math

import numpy as np
import pytest

import oracle as orc
from swraster_viewer_b200 import scenes

F32 = np.float32


def unpack(d):
    d = np.asarray(d, np.uint32)
    return np.stack([(d >> 24) & 255, (d >> 16) & 255, (d >> 8) & 255, d & 255], -1).astype(np.int32)


def cross_from_faces(faces):
    """(6, h, w, 4) uint8 in the order +X -X +Y -Y +Z -Z -> the viewer's cross image (texture.rs:923-926)."""
    _, h, w, _ = faces.shape
    img = np.zeros((3 * h, 4 * w, 4), np.uint8)
    for f, (fx, fy) in enumerate([(2, 1), (0, 1), (1, 0), (1, 2), (1, 1), (3, 1)]):
        img[fy * h:(fy + 1) * h, fx * w:(fx + 1) * w] = faces[f]
    return img


def s2l(x):
    return np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


# ---- float64 restatement of the sampling chain (texture.rs:135-165, 237-287) -----------------------------------------
def radical_inverse(i):
    b = int(i)
    b = ((b << 16) | (b >> 16)) & 0xFFFFFFFF
    b = (((b & 0x55555555) << 1) | ((b & 0xAAAAAAAA) >> 1)) & 0xFFFFFFFF
    b = (((b & 0x33333333) << 2) | ((b & 0xCCCCCCCC) >> 2)) & 0xFFFFFFFF
    b = (((b & 0x0F0F0F0F) << 4) | ((b & 0xF0F0F0F0) >> 4)) & 0xFFFFFFFF
    b = (((b & 0x00FF00FF) << 8) | ((b & 0xFF00FF00) >> 8)) & 0xFFFFFFFF
    return b * 2.3283064e-10


def norm(v):
    return v / np.linalg.norm(v)


def ggx_sample(xi, n, rough):
    a = rough * rough
    phi = 2 * math.pi * xi[0]
    ct = math.sqrt((1 - xi[1]) / (1 + (a * a - 1) * xi[1]))
    st = math.sqrt(max(1 - ct * ct, 0))
    up = np.array([0, 0, 1.0]) if abs(n[2]) < 0.999 else np.array([1.0, 0, 0])
    t = norm(np.cross(n, up))
    b = np.cross(n, t)
    return norm(t * (math.cos(phi) * st) + b * (math.sin(phi) * st) + n * ct)


def face_dir(face, u, v):
    return norm(np.array([[1, -v, -u], [-1, -v, u], [u, 1, v], [u, -1, -v], [u, -v, 1], [-u, -v, -1]][face], np.float64))


def dir_face_uv(n):
    ax, ay, az = abs(n[0]), abs(n[1]), abs(n[2])
    if ax >= ay and ax >= az:
        return (0, -n[2] / ax * .5 + .5, -n[1] / ax * .5 + .5) if n[0] >= 0 else (1, n[2] / ax * .5 + .5, -n[1] / ax * .5 + .5)
    if ay > ax and ay >= az:
        return (2, n[0] / ay * .5 + .5, n[2] / ay * .5 + .5) if n[1] >= 0 else (3, n[0] / ay * .5 + .5, -n[2] / ay * .5 + .5)
    return (4, n[0] / az * .5 + .5, -n[1] / az * .5 + .5) if n[2] >= 0 else (5, -n[0] / az * .5 + .5, -n[1] / az * .5 + .5)


def sample_linear(faces_f, d):
    """faces_f: (6, h, w, 3) sRGB in [0,1]; bilinear with clamp, then sRGB -> linear (texture.rs:274-287)."""
    f, u, v = dir_face_uv(d)
    h, w = faces_f.shape[1:3]
    u, v = min(max(u, 0), 1), min(max(v, 0), 1)
    xf, yf = u * w - .5, v * h - .5
    x0, y0 = math.floor(xf), math.floor(yf)
    fx, fy = xf - x0, yf - y0
    cl = lambda t, n: int(min(max(t, 0), n - 1))
    x0i, x1i, y0i, y1i = cl(x0, w), cl(x0 + 1, w), cl(y0, h), cl(y0 + 1, h)
    c = (faces_f[f, y0i, x0i] * (1 - fx) * (1 - fy) + faces_f[f, y0i, x1i] * fx * (1 - fy) + faces_f[f, y1i, x0i] * (1 - fx) * fy + faces_f[f, y1i, x1i] * fx * fy)
    return s2l(c)


# ---------------------------------------------------------------------------------------------------------------------
def test_brdf_lut_against_the_float64_restatement_and_limits():
    assert orc.integrate_brdf(1.0, 1e-4) == (1.0, 0.0)  # mirror, head on: all energy in the scale term
    a, b = orc.integrate_brdf(0.5, 0.5)
    assert 0 < b < a < 1 and a + b <= 1.0
    N = 16
    got = orc.bake_brdf_lut(N).reshape(-1)
    # float64 restatement of integrate_brdf / generate_brdf_lut (texture.rs:167-235) INCLUDING the tangent frame of
    # importance_sample_ggx: with n = +Z the frame is (Y, -X, Z), and the 128-sample sum is not rotation invariant
    # (scenes.brdf_lut, which builds the test inputs, ignores the frame and differs by up to 8 LSB)
    want = np.zeros((N, N, 4), np.int32)
    nz = np.array([0.0, 0.0, 1.0])
    for yy in range(N):
        rough = max(min((yy + 0.5) / N, 1.0), 1e-4)
        for xx in range(N):
            ndv = max(min((xx + 0.5) / N, 1.0), 1e-4)
            v = np.array([math.sqrt(max(1 - ndv * ndv, 0)), 0.0, ndv])
            A = B = 0.0
            for i in range(128):
                h = ggx_sample((i / 128, radical_inverse(i)), nz, rough)
                l = norm(h * (2 * np.dot(v, h)) - v)
                ndl, ndh, vdh = max(l[2], 0), max(h[2], 0), max(np.dot(v, h), 0)
                if ndl > 0:
                    k = (rough * rough + 1) ** 2 * 0.125
                    gv, gl = ndv / (ndv * (1 - k) + k), ndl / (ndl * (1 - k) + k)
                    gvis = max(gv * gl * vdh / (ndh * max(ndv, 1e-5)), 0)
                    fc = (1 - vdh) ** 5
                    A += (1 - fc) * gvis
                    B += fc * gvis
            want[yy, xx] = [int(min(max(A / 128, 0), 1) * 255), int(min(max(B / 128, 0), 1) * 255), 0, 255]
    d = np.abs(unpack(got) - want.reshape(-1, 4))
    assert d.max() <= 1 and np.count_nonzero(d) <= 8, (d.max(), np.count_nonzero(d))  # f32 vs f64 evaluation at 8-bit boundaries
    assert (unpack(got)[:, 2] == 0).all() and (unpack(got)[:, 3] == 255).all()


def test_irradiance_sh_of_a_constant_sky_and_against_the_restatement():
    c = 180
    sh = orc.bake_irradiance_sh4(scenes.pack_rgba8(np.full((6, 16, 16, 4), c, np.uint8)))
    L = float(s2l(c / 255.0))
    # constant radiance L: sum of the texel solid angles is 4 pi, and 4 pi * 0.282095^2 = 1, so coefficient 0 is pi * L; the rest vanish
    assert np.allclose(sh[0], math.pi * L, rtol=2e-3)
    assert np.abs(sh[1:]).max() < 1e-4
    # a smooth sky: against the float64 nearest-texel restatement in scenes.py (bilinear at a texel centre returns that texel)
    tex, col = scenes.sky_cubemap(16, 5)
    faces = unpack(tex.data[:6 * 256]).reshape(6, 16, 16, 4).astype(np.uint8)
    sh = orc.bake_irradiance_sh4(scenes.pack_rgba8(faces))
    want = scenes.irradiance_sh4(faces[..., :3].astype(np.float32) / np.float32(255.0))
    assert np.allclose(sh, want, rtol=2e-4, atol=2e-5), (sh, want)


def test_prefiltered_cubemap_against_the_restatement():
    rng = np.random.default_rng(9)
    faces = rng.integers(0, 256, (6, 8, 8, 4), dtype=np.uint8)
    faces[..., 3] = 255
    S = 16
    pre = orc.bake_prefilter_specular(scenes.pack_rgba8(faces), S)
    nm = pre.shape[0]
    data, offs, st = pre.reshape(-1), [m * 384 for m in range(nm)], [64] * nm
    faces_f = faces[..., :3].astype(np.float64) / 255.0
    t = ((np.arange(8) + 0.5) / 8) * 2 - 1
    worst = 0
    for mip, face, y, x in [(0, 0, 0, 0), (0, 3, 5, 2), (1, 1, 3, 4), (1, 4, 7, 7), (2, 2, 0, 6), (3, 5, 4, 1), (3, 0, 2, 2), (2, 3, 6, 0)]:
        r = face_dir(face, t[x], t[y])
        rough = mip / (nm - 1)
        if mip == 0:
            col = sample_linear(faces_f, r)
        else:
            acc, tot = np.zeros(3), 0.0
            for i in range(S):
                h = ggx_sample((i / S, radical_inverse(i)), r, max(rough, 0.045))
                l = norm(h * (2 * np.dot(r, h)) - r)
                ndl = max(np.dot(r, l), 0)
                if ndl > 0:
                    acc += sample_linear(faces_f, l) * ndl
                    tot += ndl
            col = acc / tot if tot > 0 else sample_linear(faces_f, r)
        want = np.floor(np.append(col, 1.0) * 255.0).astype(np.int32)
        got = unpack(data[offs[mip] + face * st[mip] + y * 8 + x])
        worst = max(worst, np.abs(got - want).max())
        assert np.abs(got - want).max() <= 1, (mip, face, y, x, got, want)
    # every prefiltered mip is much smoother than the sharp mip 0 (noise input, 16 samples: no finer statement holds)
    spread = [unpack(data[offs[m]:offs[m] + 384])[:, :3].std() for m in range(nm)]
    assert max(spread[1:]) < 0.5 * spread[0], spread
    # a constant sky stays constant at every roughness
    cdata = unpack(orc.bake_prefilter_specular(scenes.pack_rgba8(np.full((6, 4, 4, 4), 90, np.uint8)), 8).reshape(-1))
    lin = int(math.floor(float(s2l(90 / 255.0)) * 255.0))
    assert np.abs(cdata[:, :3] - lin).max() <= 1 and (cdata[:, 3] == 255).all()

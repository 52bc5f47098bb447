"""JPEG decoder of the loader (host/swr_jpeg.hpp) against libjpeg-turbo (through Pillow), bit for bit: baseline and
progressive, grey and colour, 4:4:4 / 4:2:2 / 4:2:0 (triangle-filter chroma upsampling), odd sizes, restart intervals,
optimised Huffman tables. CPU only."""
import io

import numpy as np
import pytest
from PIL import Image

from swraster_viewer_b200 import gltf


def make_image(w, h, seed, smooth=True):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(x / 7.0 + seed) * np.cos(y / 5.0), 128 + 90 * np.cos(x / 3.0) * np.sin(y / 11.0 + 1), (x * 3 + y * 5) % 256], -1)
    noise = rng.normal(0, 6 if smooth else 60, (h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def encode(arr, mode="RGB", **kw):
    img = Image.fromarray(arr if mode == "RGB" else arr[..., 0], mode)
    buf = io.BytesIO()
    img.save(buf, "JPEG", **kw)
    data = buf.getvalue()
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
    return data, want


CASES = [
    dict(mode="L", quality=90),
    dict(mode="L", quality=35, progressive=True),
    dict(mode="RGB", quality=92, subsampling=0),
    dict(mode="RGB", quality=75, subsampling=1),
    dict(mode="RGB", quality=75, subsampling=2),
    dict(mode="RGB", quality=50, subsampling=2, optimize=True),
    dict(mode="RGB", quality=85, subsampling=0, progressive=True),
    dict(mode="RGB", quality=60, subsampling=1, progressive=True),
    dict(mode="RGB", quality=80, subsampling=2, progressive=True),
    dict(mode="RGB", quality=100, subsampling=2),
    dict(mode="RGB", quality=5, subsampling=2),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("size", [(64, 48), (37, 29), (8, 8), (1, 1), (130, 3)])
def test_bit_exact_against_libjpeg_turbo(case, size):
    kw = dict(CASES[case])
    mode = kw.pop("mode")
    arr = make_image(size[0], size[1], case * 7 + size[0], smooth=case % 2 == 0)
    data, want = encode(arr, mode, **kw)
    got = gltf.decode_png(data)  # PNG or JPEG by signature
    assert got.shape == want.shape
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() == 0, (kw, size, int(d.max()), int(np.count_nonzero(d)))


@pytest.mark.parametrize("kw", [dict(quality=80, subsampling=2, restart_marker_blocks=3), dict(quality=70, subsampling=0, restart_marker_rows=1),
                                dict(quality=85, subsampling=1, progressive=True, restart_marker_blocks=5),
                                dict(quality=90, subsampling=2, progressive=True, optimize=True, restart_marker_rows=2)])
def test_restart_intervals(kw):
    arr = make_image(150, 70, 3, smooth=False)
    data, want = encode(arr, "RGB", **kw)
    assert b"\xff\xdd" in data  # a DRI segment is really there
    assert np.array_equal(gltf.decode_png(data), want)


def test_sixteen_bit_quantisation_tables_and_large_image():
    arr = make_image(517, 389, 11, smooth=False)
    q = [min(16 + 12 * i, 1000) for i in range(64)]  # values above 255 force 16-bit table entries (Pq = 1)
    data, want = encode(arr, "RGB", qtables=[q, q], subsampling=2)
    assert np.array_equal(gltf.decode_png(data), want)
    data, want = encode(arr, "RGB", quality=95, subsampling=1, progressive=True)
    assert np.array_equal(gltf.decode_png(data), want)


def test_damaged_files_are_rejected_or_decoded_but_never_crash():
    arr = make_image(64, 40, 5)
    rng = np.random.default_rng(8)
    outcomes = [0, 0]
    for prog in (False, True):
        data, _ = encode(arr, "RGB", quality=80, subsampling=2, progressive=prog)
        for it in range(150):
            d = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(2, len(d)))
                op = int(rng.integers(0, 3))
                if op == 0:
                    d[pos] = int(rng.integers(0, 256))
                elif op == 1:
                    del d[pos:pos + int(rng.integers(1, 40))]
                else:
                    d[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 8)), dtype=np.uint8))
            try:
                out = gltf.decode_png(bytes(d))
                assert out.ndim == 3 and out.shape[2] == 4
                outcomes[0] += 1
            except gltf.GltfError:
                outcomes[1] += 1
    assert sum(outcomes) == 300
    with pytest.raises(gltf.GltfError, match="JPEG"):
        gltf.decode_png(b"\xff\xd8\xff\xe0\x00\x10JFIF" + b"\0" * 40)


def test_jpeg_texture_through_the_loader(tmp_path):
    import json
    from test_gltf_loader import tri_doc, write_gltf, tex_arrays
    from swraster_viewer_b200 import scenes
    arr = make_image(32, 16, 2)
    Image.fromarray(arr, "RGB").save(tmp_path / "albedo.jpg", quality=88)
    want = np.asarray(Image.open(tmp_path / "albedo.jpg").convert("RGBA"))
    doc = tri_doc(extra_root={"images": [{"uri": "albedo.jpg"}], "textures": [{"source": 0}],
                              "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}]})
    g = gltf.load_gltf(write_gltf(tmp_path, doc))
    t = g.desc().textures[0]
    assert (t.width, t.height, t.max_mip_level) == (32, 16, 5)
    assert np.array_equal(tex_arrays(t)[0][:512], scenes.pack_rgba8(want).reshape(-1))


# ---- a tiny baseline encoder with arbitrary sampling factors (Pillow only writes 4:4:4 / 4:2:2 / 4:2:0) -----------------
STD_DC_L = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
STD_AC_L = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d],
            [0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42,
             0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35,
             0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67,
             0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98,
             0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7,
             0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4,
             0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa])
ZZ = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15,
      23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def huff_codes(bits, vals):
    codes, code, k = {}, 0, 0
    for l in range(1, 17):
        for _ in range(bits[l - 1]):
            codes[vals[k]] = (code, l)
            code += 1
            k += 1
        code <<= 1
    return codes


def tiny_encode(planes, factors, W, H, q=12):
    """planes: full-resolution (H, W) uint8 per component; factors: [(h, v)] per component. Box-filter downsampling,
    float DCT, one flat quantisation table, the standard luminance Huffman tables for everything, interleaved scan."""
    hmax, vmax = max(f[0] for f in factors), max(f[1] for f in factors)
    mx, my = -(-W // (8 * hmax)), -(-H // (8 * vmax))
    dcc, acc = huff_codes(*STD_DC_L), huff_codes(*STD_AC_L)
    k = np.arange(8)
    D = np.sqrt(2 / 8) * np.cos((2 * k[None] + 1) * k[:, None] * np.pi / 16)
    D[0] /= np.sqrt(2)
    comps = []
    for p, (h, v) in zip(planes, factors):
        hr, vr = hmax // h, vmax // v
        cw, ch = -(-W * h // hmax), -(-H * v // vmax)
        pad = np.pad(p.astype(np.float64), ((0, ch * vr - H), (0, cw * hr - W)), mode="edge")
        small = pad.reshape(ch, vr, cw, hr).mean(axis=(1, 3))
        full = np.pad(small, ((0, my * v * 8 - ch), (0, mx * h * 8 - cw)), mode="edge") - 128.0
        comps.append(full)
    bits = []
    pred = [0] * len(planes)

    def put(code, length):
        bits.append(format(code, f"0{length}b") if length else "")

    def category(vv):
        a = abs(vv)
        return a.bit_length()

    for yy in range(my):
        for xx in range(mx):
            for ci, (h, v) in enumerate(factors):
                for bv in range(v):
                    for bh in range(h):
                        y0, x0 = (yy * v + bv) * 8, (xx * h + bh) * 8
                        blk = D @ comps[ci][y0:y0 + 8, x0:x0 + 8] @ D.T
                        qz = np.round(blk / q).astype(int).reshape(-1)
                        zz = [int(qz[ZZ[i]]) for i in range(64)]
                        diff = zz[0] - pred[ci]
                        pred[ci] = zz[0]
                        c = category(diff)
                        put(*dcc[c])
                        if c:
                            put(diff if diff > 0 else diff + (1 << c) - 1, c)
                        run = 0
                        last = max([i for i in range(1, 64) if zz[i]], default=0)
                        for i in range(1, last + 1):
                            if zz[i] == 0:
                                run += 1
                                continue
                            while run > 15:
                                put(*acc[0xF0])
                                run -= 16
                            c = category(zz[i])
                            put(*acc[(run << 4) | c])
                            put(zz[i] if zz[i] > 0 else zz[i] + (1 << c) - 1, c)
                            run = 0
                        if last < 63:
                            put(*acc[0x00])
    s = "".join(bits)
    s += "1" * (-len(s) % 8)
    ecs = bytearray()
    for i in range(0, len(s), 8):
        b = int(s[i:i + 8], 2)
        ecs.append(b)
        if b == 0xFF:
            ecs.append(0)

    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload
    out = b"\xff\xd8" + seg(0xE0, b"JFIF\0\x01\x01\0\0\x01\0\x01\0\0") + seg(0xDB, bytes([0]) + bytes([q] * 64))
    out += seg(0xC0, bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([len(planes)]) +
               b"".join(bytes([i + 1, (h << 4) | v, 0]) for i, (h, v) in enumerate(factors)))
    out += seg(0xC4, bytes([0x00]) + bytes(STD_DC_L[0]) + bytes(STD_DC_L[1]) + bytes([0x10]) + bytes(STD_AC_L[0]) + bytes(STD_AC_L[1]))
    out += seg(0xDA, bytes([len(planes)]) + b"".join(bytes([i + 1, 0x00]) for i in range(len(planes))) + bytes([0, 63, 0]))
    return out + bytes(ecs) + b"\xff\xd9"


@pytest.mark.parametrize("factors", [[(1, 2), (1, 1), (1, 1)],   # 4:4:0: vertical triangle filter (h1v2)
                                     [(4, 1), (1, 1), (1, 1)],   # 4:1:1: replication
                                     [(2, 2), (2, 1), (1, 2)],   # mixed ratios per component
                                     [(2, 2), (1, 1), (1, 1)],   # 4:2:0 through this encoder as a control
                                     [(1, 1), (1, 1), (1, 1)]])
@pytest.mark.parametrize("size", [(40, 24), (29, 19)])
def test_other_sampling_factors_against_libjpeg_turbo(factors, size):
    W, H = size
    arr = make_image(W, H, W + len(factors))
    data = tiny_encode([arr[..., 0], arr[..., 1], arr[..., 2]], factors, W, H)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
    got = gltf.decode_png(data)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() == 0, (factors, size, int(d.max()), int(np.count_nonzero(d)))

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

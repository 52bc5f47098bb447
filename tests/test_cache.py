"""The reference's bake caches (`.ggx` texture.rs:12-17, 422-552; `.gi` gi.rs:17-22, 30-122) through the host
library (include/swr_gltf.h swrh_ggx_cache_* / swrh_gi_cache_*). Checked against the file layout the reference writes — header
bytes restated here, payload handled by the SYSTEM's brotli through ctypes (an independent producer / consumer of the stream:
other quality, other window) — and against the reference's "that is not my cache" rules."""
import ctypes as C
import ctypes.util
import os
import struct

import numpy as np
import pytest

from swraster_viewer_b200 import gltf


def _brotli():
    try:
        dec, enc = C.CDLL("libbrotlidec.so.1"), C.CDLL("libbrotlienc.so.1")
    except OSError:
        pytest.skip("the system has no libbrotli: the cache calls fail loudly there (tested below)")
    dec.BrotliDecoderDecompress.argtypes = [C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
    enc.BrotliEncoderCompress.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
    enc.BrotliEncoderMaxCompressedSize.restype = C.c_size_t
    enc.BrotliEncoderMaxCompressedSize.argtypes = [C.c_size_t]
    return dec, enc


def inflate(stream, nbytes):
    dec, _ = _brotli()
    out = C.create_string_buffer(nbytes + 16)
    n = C.c_size_t(nbytes + 16)
    assert dec.BrotliDecoderDecompress(len(stream), stream, C.byref(n), out) == 1
    return out.raw[:n.value]


def deflate(raw, quality=9, lgwin=18):
    _, enc = _brotli()
    cap = enc.BrotliEncoderMaxCompressedSize(len(raw)) + 64
    out = C.create_string_buffer(cap)
    n = C.c_size_t(cap)
    assert enc.BrotliEncoderCompress(quality, lgwin, 0, len(raw), raw, C.byref(n), out)
    return out.raw[:n.value]


def test_ggx_cache_file_layout_and_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    w, h = 16, 8
    mips = 5  # 1 + ilog2(16)
    tex = rng.integers(0, 2**32, size=(mips, 6, h, w), dtype=np.uint64).astype(np.uint32)
    tex[1:] &= 0xF0F0F0FF  # something for the compressor to find
    p = tmp_path / "cubemap.ggx"
    gltf.ggx_cache_save(p, tex)
    blob = p.read_bytes()
    assert blob[:4] == b"GGX0" and struct.unpack("<4I", blob[4:20]) == (1, w, h, mips)  # texture.rs:527-533
    assert inflate(blob[20:], tex.size * 4) == tex.astype("<u4").tobytes()                # texture.rs:535-538
    back = gltf.ggx_cache_load(p, w, h)
    assert back is not None and np.array_equal(back, tex)
    # a cache written by somebody else's encoder settings reads the same (the reference only fixes the format, not the stream)
    q = tmp_path / "other.ggx"
    q.write_bytes(b"GGX0" + struct.pack("<4I", 1, w, h, mips) + deflate(tex.astype("<u4").tobytes()))
    assert np.array_equal(gltf.ggx_cache_load(q, w, h), tex)


def test_ggx_cache_that_is_not_mine(tmp_path):
    w, h, mips = 8, 8, 4
    tex = np.arange(mips * 6 * h * w, dtype=np.uint32).reshape(mips, 6, h, w)
    raw = tex.astype("<u4").tobytes()
    good = b"GGX0" + struct.pack("<4I", 1, w, h, mips) + deflate(raw)
    cases = {
        "missing": None,
        "short": good[:12],
        "magic": b"GGX1" + good[4:],
        "version": good[:4] + struct.pack("<I", 2) + good[8:],
        "size": good[:8] + struct.pack("<2I", w * 2, h) + good[16:],
        "mips": good[:16] + struct.pack("<I", mips + 1) + good[20:],
        "payload_short": b"GGX0" + struct.pack("<4I", 1, w, h, mips) + deflate(raw[:-4]),
        "payload_long": b"GGX0" + struct.pack("<4I", 1, w, h, mips) + deflate(raw + b"\0\0\0\0"),
        "truncated_stream": good[:-7],
    }
    for name, blob in cases.items():
        p = tmp_path / f"{name}.ggx"
        if blob is not None:
            p.write_bytes(blob)
        assert gltf.ggx_cache_load(p, w, h) is None, name  # texture.rs:431-470: Ok(None) -> the caller bakes
    p = tmp_path / "good.ggx"
    p.write_bytes(good)
    assert np.array_equal(gltf.ggx_cache_load(p, w, h), tex)
    assert gltf.ggx_cache_load(p, w, h * 2) is None  # expected size comes from the sky that was loaded (scene.rs:165-169)


def test_gi_cache_round_trip_keeps_every_bit(tmp_path):
    rng = np.random.default_rng(2)
    dims = (5, 3, 4)
    n = dims[0] * dims[1] * dims[2]
    gi = rng.standard_normal((n, 4, 4)).astype(np.float32)
    gi.view(np.uint32)[0, 0, :] = [0x7FC00001, 0x7F800000, 0x80000000, 0x00000001]  # NaN payload, inf, -0, denormal
    p = tmp_path / "scene.gi"
    gltf.gi_cache_save(p, dims, gi)
    blob = p.read_bytes()
    assert blob[:4] == b"VGI0" and struct.unpack("<4I", blob[4:20]) == (8, *dims)       # gi.rs:17-18, 87-93
    assert inflate(blob[20:], gi.size * 4) == gi.astype("<f4").tobytes()                # voxel-major, coefficient, then r g b w
    back = gltf.gi_cache_load(p, dims)
    assert back is not None and np.array_equal(back.view(np.uint32), gi.view(np.uint32))
    assert gltf.gi_cache_load(p, (5, 3, 5)) is None and gltf.gi_cache_load(tmp_path / "none.gi", dims) is None
    old = tmp_path / "v7.gi"
    old.write_bytes(blob[:4] + struct.pack("<I", 7) + blob[8:])
    assert gltf.gi_cache_load(old, dims) is None  # GI_CACHE_VERSION = 8 only


def test_save_rejects_a_texture_of_the_wrong_size(tmp_path):
    host = gltf._host()
    host.swrh_ggx_cache_save.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    assert host.swrh_ggx_cache_save(None, 4, 4, 3, None) != 0  # NULL arguments are errors, not crashes
    with pytest.raises(gltf.GltfError):
        gltf.ggx_cache_save(tmp_path / "no" / "such" / "dir" / "x.ggx", np.zeros((3, 6, 4, 4), np.uint32))


def test_mutated_caches_never_crash_the_reader(tmp_path):
    """Hostile / damaged files: every byte-flipped or truncated cache is either rejected ("no cache") or decodes to a payload of
    exactly the expected size — never a crash, never an error the reference would not raise (its decoder errors are I/O errors,
    ours are folded into "no cache" because the one-shot decoder does not tell them from a short stream)."""
    rng = np.random.default_rng(7)
    w, h, mips = 8, 4, 4
    tex = rng.integers(0, 2**32, size=(mips, 6, h, w), dtype=np.uint64).astype(np.uint32)
    good = tmp_path / "good.ggx"
    gltf.ggx_cache_save(good, tex)
    blob = bytearray(good.read_bytes())
    p = tmp_path / "m.ggx"
    accepted = 0
    for trial in range(300):
        b = bytearray(blob)
        kind = trial % 3
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            i = int(rng.integers(20, len(b)))
            b[i:i] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
        p.write_bytes(bytes(b))
        got = gltf.ggx_cache_load(p, w, h)
        if got is not None:
            assert got.shape == tex.shape
            accepted += 1
    assert accepted < 300  # most damage is caught by the header checks, the brotli decoder or the exact-length rule
    gi = rng.standard_normal((2 * 2 * 2, 4, 4)).astype(np.float32)
    gltf.gi_cache_save(tmp_path / "g.gi", (2, 2, 2), gi)
    gb = bytearray((tmp_path / "g.gi").read_bytes())
    for trial in range(100):
        b = bytearray(gb)
        b[int(rng.integers(0, len(b)))] ^= 0xFF
        (tmp_path / "mg.gi").write_bytes(bytes(b[:int(rng.integers(1, len(b) + 1))]))
        got = gltf.gi_cache_load(tmp_path / "mg.gi", (2, 2, 2))
        assert got is None or got.shape == gi.shape

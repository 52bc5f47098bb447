"""Host-side property tests of two small pieces of device arithmetic that decide WORK, not results, but whose mistakes would be
hard to see on the GPU: the 32-bit fast path of the exactness bound (csrc/swr_device.cuh edge_bound_ok: picks between the integer
edge evaluation and the f32 chain replay) and the depth bucket of the tile lists (csrc/swr_raster.cuh depth_bucket). The
functions are cut out of the CUDA sources, compiled for the host with a few shims and compared with exact Python arithmetic."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "swraster-viewer_b200", "csrc")


def cut(path, signature):
    src = open(path).read()
    i = src.index(signature)
    depth, j = 0, src.index("{", i)
    for k in range(j, len(src)):
        depth += {"{": 1, "}": -1}.get(src[k], 0)
        if depth == 0:
            return src[i:k + 1]
    raise AssertionError(signature)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("devmath")
    raster = open(os.path.join(CSRC, "swr_raster.cuh")).read()
    shift = re.search(r"#define SWR_ZBUCKET_SHIFT (\d+)", raster).group(1)
    code = """
#include <cstdint>
#include <cstring>
#include <cmath>
#include <algorithm>
#define __device__
#define __forceinline__ inline
using std::min;
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
#define SWR_ZBUCKET_SHIFT %s
#define SWR_ZDEPTHS (8 << SWR_ZBUCKET_SHIFT)
#define SWR_ZBUCKETS SWR_ZDEPTHS
%s
%s
extern "C" int t_edge_bound_ok(int a, int b, int c, uint32_t xhi, uint32_t yhi) { return edge_bound_ok(a, b, c, xhi, yhi) ? 1 : 0; }
extern "C" uint32_t t_depth_bucket(float z) { return depth_bucket(z); }
extern "C" uint32_t t_nbuckets() { return SWR_ZBUCKETS; }
#define SWR_TILE 64
struct uint4 { uint32_t x, y, z, w; };
%s
%s
%s
%s
extern "C" int t_key_index(int x, int y) { return key_index(x, y); }
extern "C" uint32_t t_counts_to_cursors(const uint32_t *count, uint32_t *cursor, int tile, uint32_t base) {
    uint32_t c[SWR_ZBUCKETS];
    const uint32_t total = load_tile_counts(count, tile, c);
    store_tile_cursors(cursor, tile, base, c);
    return total;
}
""" % (shift, cut(os.path.join(CSRC, "swr_device.cuh"), "__device__ __forceinline__ bool edge_bound_ok("),
       cut(os.path.join(CSRC, "swr_raster.cuh"), "__device__ __forceinline__ uint32_t depth_bucket("),
       cut(os.path.join(CSRC, "swr_raster.cuh"), "__device__ __forceinline__ int key_swz("),
       cut(os.path.join(CSRC, "swr_raster.cuh"), "__device__ __forceinline__ int key_index("),
       cut(os.path.join(CSRC, "swr_raster.cuh"), "__device__ __forceinline__ uint32_t load_tile_counts("),
       cut(os.path.join(CSRC, "swr_raster.cuh"), "__device__ __forceinline__ void store_tile_cursors("))
    src = d / "devmath.cpp"
    src.write_text(code)
    so = d / "devmath.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)])
    lib = C.CDLL(str(so))
    lib.t_edge_bound_ok.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32]
    lib.t_depth_bucket.argtypes = [C.c_float]
    lib.t_depth_bucket.restype = C.c_uint32
    lib.t_nbuckets.restype = C.c_uint32
    lib.t_counts_to_cursors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]
    lib.t_counts_to_cursors.restype = C.c_uint32
    return lib


def test_exactness_bound_fast_path_equals_the_wide_formula(lib):
    rng = np.random.default_rng(11)
    lim = 1 << 24
    cases = []
    # around the 2^12 switch, around the 2^24 limit, extremes of i32, and plain random triples
    for _ in range(20000):
        scale = int(rng.choice([1 << 4, 1 << 8, 1 << 11, 1 << 12, 1 << 13, 1 << 20, 1 << 31]))
        a, b = (int(rng.integers(-scale, scale)) for _ in range(2))
        c = int(rng.integers(-(1 << int(rng.integers(1, 32))), 1 << int(rng.integers(1, 32))))
        xhi, yhi = int(rng.integers(0, 261377)), int(rng.integers(0, 261377))
        cases.append((a, b, c, xhi, yhi))
    for v in (-2**31, 2**31 - 1, 4095, 4096, -4096, -4095, 0):
        for c in (-2**31, 2**24 - 1, 2**24, -(2**24), 0):
            cases += [(v, 3, c, 1024, 1024), (3, v, c, 261376, 0), (v, v, c, 0, 0)]
    # boundary: pick c so that the sum lands exactly on / next to 2^24
    for _ in range(5000):
        a, b = int(rng.integers(-4095, 4096)), int(rng.integers(-4095, 4096))
        xhi, yhi = int(rng.integers(0, 4000)), int(rng.integers(0, 4000))
        rest = lim - abs(a) * xhi - abs(b) * yhi
        for dc in (-1, 0, 1):
            c = rest + dc
            if -2**31 <= c < 2**31:
                cases.append((a, b, c, xhi, yhi))
                cases.append((a, b, -c if -c >= -2**31 and -c < 2**31 else c, xhi, yhi))
    for a, b, c, xhi, yhi in cases:
        a, b, c = max(min(a, 2**31 - 1), -2**31), max(min(b, 2**31 - 1), -2**31), max(min(c, 2**31 - 1), -2**31)
        want = abs(a) * xhi + abs(b) * yhi + abs(c) < lim
        assert bool(lib.t_edge_bound_ok(a, b, c, xhi, yhi)) == want, (a, b, c, xhi, yhi)


def test_depth_bucket_is_monotonic_and_covers_the_range(lib):
    nb = lib.t_nbuckets()
    zs = np.concatenate([np.linspace(-2.0, 1.0, 20001), 1.0 - np.logspace(-9, 0, 2000), [np.inf, -np.inf]]).astype(np.float32)
    zs = np.sort(zs[np.isfinite(zs)])
    b = np.array([lib.t_depth_bucket(float(z)) for z in zs])
    assert (np.diff(b) >= 0).all(), "a deeper triangle must never land in a nearer bucket"
    assert b.min() == 0 and b.max() == nb - 1
    assert lib.t_depth_bucket(0.0) == 0 and lib.t_depth_bucket(-5.0) == 0          # at / in front of the near plane
    assert lib.t_depth_bucket(1.0) == nb - 1 and lib.t_depth_bucket(7.0) == nb - 1  # far plane and beyond
    assert lib.t_depth_bucket(float("nan")) == nb - 1 and lib.t_depth_bucket(float("inf")) == nb - 1
    # octaves of 1 - z: z = 1 - 2^-k sits in bucket k (with SWR_ZBUCKET_SHIFT = 0)
    if nb == 8:
        for k in range(0, 7):
            assert lib.t_depth_bucket(float(np.float32(1.0) - np.float32(2.0 ** -k))) == min(k, nb - 1)


def test_key_swizzle_is_a_bijection_that_keeps_quad_rows_paired(lib):
    """Shared-memory / global key layout of a tile: every pixel has its own slot, a row stays inside its 64 slots, and the two pixels
    of a quad row (even x, x + 1) stay an aligned pair (one 16-byte access) — k_raster_tiles and load_key rely on all three."""
    idx = np.array([[lib.t_key_index(x, y) for x in range(64)] for y in range(64)])
    assert sorted(idx.ravel().tolist()) == list(range(4096))
    assert (idx // 64 == np.arange(64)[:, None]).all()
    assert (idx[:, 0::2] % 2 == 0).all() and (idx[:, 1::2] == idx[:, 0::2] + 1).all()
    # rows two apart map the same x to different banks (what the swizzle is for): 8 consecutive quad rows, 8 distinct offsets
    for x in (0, 2, 30):
        assert len({int(idx[y, x]) % 64 for y in range(0, 16, 2)}) == 8


def test_bucket_counters_become_consecutive_cursors(lib):
    nb = lib.t_nbuckets()
    rng = np.random.default_rng(3)
    ntiles = 37
    count = rng.integers(0, 5000, size=ntiles * nb, dtype=np.uint32)
    cursor = np.zeros_like(count)
    base = 0
    for t in range(ntiles):
        total = lib.t_counts_to_cursors(count.ctypes.data, cursor.ctypes.data, t, base)
        c = count[t * nb:(t + 1) * nb].astype(np.int64)
        assert total == c.sum()
        assert np.array_equal(cursor[t * nb:(t + 1) * nb], base + np.concatenate([[0], np.cumsum(c)[:-1]]))
        base += int(total)

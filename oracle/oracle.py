"""ctypes binding of the CPU oracle (oracle/oracle.cpp). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OrcStats(C.Structure):
    _fields_ = [("triangles_submitted", C.c_uint64), ("vertices_submitted", C.c_uint64), ("triangles_binned", C.c_uint64),
                ("triangles_clipped", C.c_uint64), ("tile_refs", C.c_uint64), ("tile_refs_dup", C.c_uint64),
                ("ndraws", C.c_uint32), ("tiles", C.c_uint32), ("ms_clipbin", C.c_double), ("ms_raster", C.c_double),
                ("ms_resolve", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def _cpu_has_v3():
    try:
        flags = open("/proc/cpuinfo").read()
        return all(f in flags for f in (" avx2", " fma", " bmi2"))
    except OSError:
        return False


def build(force=False):
    libs = [os.path.join(_HERE, n) for n in ("liboracle.so", "liboracle_v3.so")]
    if force or not all(os.path.exists(p) for p in libs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return libs


_lib = None


def load():
    global _lib
    if _lib is None:
        base, v3 = build()
        _lib = C.CDLL(v3 if _cpu_has_v3() else base)
        vp = C.c_void_p
        _lib.orc_create.restype = vp
        _lib.orc_create.argtypes = [C.c_int, C.c_int]
        _lib.orc_destroy.argtypes = [vp]
        _lib.orc_render.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(OrcStats)]
        _lib.orc_resolve.argtypes = [vp, C.c_float, C.c_int, vp, C.POINTER(OrcStats)]
        _lib.orc_build_draws.argtypes = [vp, vp, vp, C.c_int]
        _lib.orc_set_exact_rsqrt.argtypes = [C.c_int]
        _lib.orc_update_auto_exposure.argtypes = [vp, vp, C.c_int, C.c_float]
        _lib.orc_tile_digests.argtypes = [vp, vp]
        _lib.orc_num_threads.restype = C.c_int
    return _lib


class Oracle:
    """Renderer::new(width, height) on the CPU restatement."""

    def __init__(self, width, height):
        self.lib = load()
        self.width, self.height = width, height
        self.h = self.lib.orc_create(width, height)
        self.stats = OrcStats()

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    def render(self, scene, camera_abi, nthreads=1, shade=True, fresh=True, outputs=True):
        """Returns dict(depth, seq, bary1, bary2, color, luminance). nthreads=1 is the serial (parity) schedule."""
        n = self.width * self.height
        tiles = ((self.width + 63) // 64) * ((self.height + 63) // 64)
        out = {}
        if outputs:
            out = dict(depth=np.empty(n, np.uint32), seq=np.empty(n, np.uint32), bary1=np.empty(n, np.float32),
                       bary2=np.empty(n, np.float32), color=np.zeros((self.height, self.width, 3), np.float32),
                       luminance=np.empty(tiles, np.float32))
        p = lambda k: out[k].ctypes.data if outputs else None
        rc = self.lib.orc_render(self.h, C.addressof(scene.desc()), C.addressof(camera_abi), nthreads, int(shade), int(fresh),
                                 p("depth"), p("seq"), p("bary1"), p("bary2"), p("color"), p("luminance"), C.byref(self.stats))
        assert rc == 0
        return out

    def resolve(self, exposure=2.0, nthreads=1):
        px = np.zeros(self.width * self.height, np.uint32)
        self.lib.orc_resolve(self.h, exposure, nthreads, px.ctypes.data, C.byref(self.stats))
        return px

    def tile_digests(self):
        """(ntiles, 3) uint64: visibility digest, covered lanes, colour digest of every tile after the last render
        (tools/reference_digest.rs prints the same from the Rust renderer)."""
        import numpy as np
        n = ((self.width + 63) // 64) * ((self.height + 63) // 64)
        out = np.zeros((n, 3), np.uint64)
        got = self.lib.orc_tile_digests(self.h, out.ctypes.data)
        assert got == n
        return out

    def build_draws(self, scene, camera_abi, draw_type, max_draws=1 << 20):
        n = self.lib.orc_build_draws(C.addressof(scene.desc()), C.addressof(camera_abi), None, 0)
        arr = (draw_type * max(n, 1))()
        self.lib.orc_build_draws(C.addressof(scene.desc()), C.addressof(camera_abi), C.addressof(arr), n)
        return arr, n


def set_exact_rsqrt(on):
    load().orc_set_exact_rsqrt(int(on))


def num_threads():
    return load().orc_num_threads()


def update_auto_exposure(state, center_luminance, delta_time):
    """renderer.rs:258-290 on a list of per-tile metering values; state = [auto_exposure, target, ev] (float32), returns the new state."""
    import numpy as np
    st = np.array(state, np.float32)
    lum = np.ascontiguousarray(center_luminance, np.float32)
    load().orc_update_auto_exposure(st.ctypes.data, lum.ctypes.data, len(lum), float(delta_time))
    return st


# ---- load-time bakes (oracle_bake.cpp, oracle_sunvis.cpp): the checkers of the device bakes ----------------------------------
def integrate_brdf(ndotv, roughness):
    """One texel's worth of integrate_brdf (texture.rs:167-197): (scale, bias)."""
    out = (C.c_float * 2)()
    lib = load()
    lib.orc_integrate_brdf.argtypes = [C.c_float, C.c_float, C.c_void_p]
    lib.orc_integrate_brdf(ndotv, roughness, C.addressof(out))
    return float(out[0]), float(out[1])


def bake_brdf_lut(size):
    import numpy as np
    out = np.zeros((size, size), np.uint32)
    lib = load()
    lib.orc_bake_brdf_lut.argtypes = [C.c_uint32, C.c_void_p]
    lib.orc_bake_brdf_lut(size, out.ctypes.data)
    return out


def bake_irradiance_sh4(faces):
    """faces: (6, h, w) uint32 RGBA8 (R in bits 31..24). Returns (4, 3) float32."""
    import numpy as np
    f = np.ascontiguousarray(faces, np.uint32)
    out = np.zeros(12, np.float32)
    lib = load()
    lib.orc_bake_irradiance_sh4.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.orc_bake_irradiance_sh4(f.ctypes.data, f.shape[2], f.shape[1], out.ctypes.data)
    return out.reshape(4, 3)


def bake_prefilter_specular(faces, sample_count):
    """faces: (6, h, w) uint32. Returns (num_mips, 6, h, w) uint32."""
    import numpy as np
    f = np.ascontiguousarray(faces, np.uint32)
    h, w = f.shape[1], f.shape[2]
    nm = int(max(w, h)).bit_length()
    out = np.zeros((nm, 6, h, w), np.uint32)
    lib = load()
    lib.orc_bake_prefilter_specular.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    got = lib.orc_bake_prefilter_specular(f.ctypes.data, w, h, sample_count, out.ctypes.data)
    assert got == nm
    return out


def sun_visibility(scene):
    """compute_sun_visibility + blur of a scenes.SceneData / loaded glTF (anything with .desc()): (D, H, W) float32."""
    import numpy as np
    d = scene.desc()
    w, h, dd = d.voxel_grid.dims[:]
    out = np.zeros(w * h * dd, np.float32)
    lib = load()
    lib.orc_sun_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    light = (C.c_float * 3)(*d.light_direction[:])
    rc = lib.orc_sun_visibility(C.addressof(d), C.addressof(d.voxel_grid), C.addressof(light), out.ctypes.data)
    assert rc == 0
    return out.reshape(dd, h, w)


def sunvis_active_mask(scene):
    import numpy as np
    d = scene.desc()
    w, h, dd = d.voxel_grid.dims[:]
    out = np.zeros(w * h * dd, np.uint8)
    lib = load()
    lib.orc_sunvis_active_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_sunvis_active_mask(C.addressof(d), C.addressof(d.voxel_grid), out.ctypes.data)
    return out.reshape(dd, h, w)

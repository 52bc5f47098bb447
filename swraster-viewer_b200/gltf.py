"""ctypes face of the native glTF / GLB loader (include/swr_gltf.h, host/swr_gltf.hpp): the step in front of the hot path
(SURVEY 8f N2). `load_gltf` returns a scene object that Renderer.render_scene and the oracle accept like a SceneData."""
import ctypes as C
import os

import numpy as np

from . import abi
from .renderer import load_libraries


class GltfEnv(C.Structure):
    _fields_ = [("cubemap", C.POINTER(abi.TextureDesc)), ("cubemap_specular", C.POINTER(abi.TextureDesc)), ("brdf_lut", C.POINTER(abi.TextureDesc)),
                ("voxel_grid", abi.VoxelGridDesc), ("light_direction", C.c_float * 3), ("light_color", C.c_float * 3)]


class GltfInfo(C.Structure):
    _fields_ = [("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3), ("bounds_center", C.c_float * 3), ("bounds_diagonal", C.c_float),
                ("ncameras", C.c_uint32), ("nfile_textures", C.c_uint32)]


class GltfCamera(C.Structure):
    _fields_ = [("perspective", C.c_int32), ("yfov_or_xmag", C.c_float), ("aspect_or_ymag", C.c_float), ("znear", C.c_float), ("zfar", C.c_float),
                ("transform", C.c_float * 16)]


_bound = False


def _host():
    global _bound
    _, host = load_libraries()
    if not _bound:
        vp, u32 = C.c_void_p, C.c_uint32
        host.swrh_gltf_load.restype = vp
        host.swrh_gltf_load.argtypes = [C.c_char_p, C.POINTER(GltfEnv)]
        host.swrh_gltf_free.argtypes = [vp]
        host.swrh_gltf_scene.restype = C.POINTER(abi.SceneDesc)
        host.swrh_gltf_scene.argtypes = [vp]
        host.swrh_gltf_get_info.argtypes = [vp, C.POINTER(GltfInfo)]
        host.swrh_gltf_get_camera.argtypes = [vp, u32, C.POINTER(GltfCamera)]
        host.swrh_gltf_texture_uri.restype = C.c_char_p
        host.swrh_gltf_texture_uri.argtypes = [vp, u32]
        host.swrh_gltf_register_image.argtypes = [C.c_char_p, vp, u32, u32]
        host.swrh_compute_smooth_normals.argtypes = [vp, u32, vp, u32, vp]
        host.swrh_compute_tangents.argtypes = [vp, vp, vp, u32, vp, u32, vp]
        host.swrh_build_mip_chain.argtypes = [vp, u32, u32, u32, vp, C.POINTER(u32), C.POINTER(u32), vp]
        host.swrh_decode_png.argtypes = [vp, C.c_size_t, vp, C.POINTER(u32), C.POINTER(u32)]
        host.swrh_last_error.restype = C.c_char_p
        _bound = True
    return host


class GltfError(RuntimeError):
    pass


def _check(rc, host):
    if rc != 0:
        raise GltfError(host.swrh_last_error().decode())


class GltfScene:
    """A loaded document. Owns the native object; `desc()` is the swr_scene_desc view the renderer consumes."""

    def __init__(self, handle, env_keepalive):
        self._host = _host()
        self._h = handle
        self._keep = env_keepalive
        info = GltfInfo()
        _check(self._host.swrh_gltf_get_info(self._h, C.byref(info)), self._host)
        self.bounds_min = np.array(info.bounds_min[:], np.float32)
        self.bounds_max = np.array(info.bounds_max[:], np.float32)
        self.bounds_center = np.array(info.bounds_center[:], np.float32)
        self.bounds_diagonal = np.float32(info.bounds_diagonal)
        self.nfile_textures = info.nfile_textures
        self.cameras = []
        for i in range(info.ncameras):
            c = GltfCamera()
            _check(self._host.swrh_gltf_get_camera(self._h, i, C.byref(c)), self._host)
            self.cameras.append(dict(perspective=bool(c.perspective), yfov_or_xmag=c.yfov_or_xmag, aspect_or_ymag=c.aspect_or_ymag, znear=c.znear,
                                     zfar=c.zfar, transform=np.array(c.transform[:], np.float32)))

    def desc(self):
        return self._host.swrh_gltf_scene(self._h).contents

    @property
    def total_triangles(self):
        d = self.desc()
        return sum(sum(d.primitives[p].nindices // 3 for p in range(d.meshes[n.mesh_index].first_primitive,
                                                                    d.meshes[n.mesh_index].first_primitive + d.meshes[n.mesh_index].num_primitives))
                   for n in (d.nodes[i] for i in range(d.nnodes)) if n.mesh_index >= 0)

    def texture_uri(self, slot):
        u = self._host.swrh_gltf_texture_uri(self._h, slot)
        return u.decode() if u is not None else None

    def close(self):
        if self._h:
            self._host.swrh_gltf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_gltf(path, environment=None):
    """Load `path` (.gltf or .glb). `environment`: a scenes.SceneData whose sky cubemap / prefiltered cubemap / BRDF LUT /
    voxel grid / light are attached to the loaded scene (a glTF file carries none of them)."""
    host = _host()
    env, keep = None, None
    if isinstance(environment, BakedEnvironment):
        # baked on the host from a sky image; the voxel grid spans the scene bounds, sun and colour as scene.rs:236-239
        env = GltfEnv.from_buffer_copy(environment.env)
        d = np.array([-0.2, 1.0, 0.5], np.float32)
        d = d / np.float32(np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1]) + np.float32(d[2] * d[2])))
        env.light_direction = (C.c_float * 3)(*d)
        env.light_color = (C.c_float * 3)(5.0, 5.0, 4.75)
        keep = environment
    elif environment is not None:
        sd = environment.desc()
        env = GltfEnv()
        env.cubemap = C.pointer(sd.textures[sd.cubemap])
        env.cubemap_specular = C.pointer(sd.textures[sd.cubemap_specular])
        env.brdf_lut = C.pointer(sd.textures[sd.brdf_lut])
        env.voxel_grid = sd.voxel_grid
        env.light_direction = sd.light_direction
        env.light_color = sd.light_color
        keep = (environment, sd)
    h = host.swrh_gltf_load(str(path).encode(), C.byref(env) if env is not None else None)
    if not h:
        raise GltfError(host.swrh_last_error().decode())
    return GltfScene(h, keep)


def register_image(uri, rgba_u8):
    """Hand the loader a decoded image for `uri` ((H, W, 4) uint8), e.g. a JPEG it cannot decode itself. None removes it."""
    host = _host()
    if rgba_u8 is None:
        _check(host.swrh_gltf_register_image(uri.encode(), None, 0, 0), host)
        return
    a = np.ascontiguousarray(rgba_u8, np.uint8)
    _check(host.swrh_gltf_register_image(uri.encode(), a.ctypes.data, a.shape[1], a.shape[0]), host)


def compute_smooth_normals(positions4, indices):
    host = _host()
    p = np.ascontiguousarray(positions4, np.float32)
    i = np.ascontiguousarray(indices, np.uint32)
    out = np.zeros_like(p)
    _check(host.swrh_compute_smooth_normals(p.ctypes.data, len(p), i.ctypes.data, len(i), out.ctypes.data), host)
    return out


def compute_tangents(positions4, texcoords2, normals4, indices):
    host = _host()
    p, uv, n = (np.ascontiguousarray(a, np.float32) for a in (positions4, texcoords2, normals4))
    i = np.ascontiguousarray(indices, np.uint32)
    out = np.zeros_like(p)
    _check(host.swrh_compute_tangents(p.ctypes.data, uv.ctypes.data, n.ctypes.data, len(p), i.ctypes.data, len(i), out.ctypes.data), host)
    return out


def build_mip_chain(base_u32, width, height, texture_type):
    """(data, offsets, widths, heights, strides) as Texture::generate_mipmaps builds them (texture.rs:45-128)."""
    host = _host()
    b = np.ascontiguousarray(base_u32, np.uint32).reshape(-1)
    nt, nm = C.c_uint32(), C.c_uint32()
    _check(host.swrh_build_mip_chain(b.ctypes.data, width, height, texture_type, None, C.byref(nt), C.byref(nm), None), host)
    data = np.zeros(nt.value, np.uint32)
    table = np.zeros(4 * nm.value, np.uint32)
    _check(host.swrh_build_mip_chain(b.ctypes.data, width, height, texture_type, data.ctypes.data, C.byref(nt), C.byref(nm), table.ctypes.data), host)
    t = table.reshape(4, nm.value)
    return data, t[0], t[1], t[2], t[3]


def decode_png(file_bytes):
    host = _host()
    buf = np.frombuffer(file_bytes, np.uint8)
    w, h = C.c_uint32(), C.c_uint32()
    _check(host.swrh_decode_png(buf.ctypes.data, len(buf), None, C.byref(w), C.byref(h)), host)
    out = np.zeros((h.value, w.value, 4), np.uint8)
    _check(host.swrh_decode_png(buf.ctypes.data, len(buf), out.ctypes.data, C.byref(w), C.byref(h)), host)
    return out


class BakedEnvironment:
    """Bakes of everything the viewer derives from its sky image (include/swr_gltf.h swrh_env_bake; the three integrals run on the GPU):
    sky cubemap + mips, GGX-prefiltered cubemap, irradiance SH4, BRDF LUT, SH-initialised voxel grid."""

    def __init__(self, cross_rgba_u8, lut_size=128, specular_samples=64, voxel_dim=16, irradiance_scale=0.25, sky_visibility=1.0, light_intensity=1.0,
                 ggx_cache=None):
        """ggx_cache: path of the reference's `.ggx` cache (scene.rs:164-206): read when it matches the sky's face size, else written
        after the prefiltered cubemap has been baked; `specular_from_cache` says which."""
        host = _host()
        host.swrh_env_bake_cached.restype = C.c_void_p
        host.swrh_env_bake_cached.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_char_p]
        host.swrh_env_specular_from_cache.argtypes = [C.c_void_p]
        host.swrh_env_get.argtypes = [C.c_void_p, C.POINTER(GltfEnv), C.POINTER(C.c_float * 12)]
        host.swrh_env_free.argtypes = [C.c_void_p]
        a = np.ascontiguousarray(cross_rgba_u8, np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        self._host = host
        self._h = host.swrh_env_bake_cached(a.ctypes.data, a.shape[1], a.shape[0], lut_size, specular_samples, voxel_dim, irradiance_scale, sky_visibility,
                                            light_intensity, os.fsencode(ggx_cache) if ggx_cache else None)
        if not self._h:
            raise GltfError(host.swrh_last_error().decode())
        self.specular_from_cache = bool(host.swrh_env_specular_from_cache(self._h))
        self.env = GltfEnv()
        sh = (C.c_float * 12)()
        _check(host.swrh_env_get(self._h, C.byref(self.env), C.byref(sh)), host)
        self.irradiance_sh = np.array(sh[:], np.float32).reshape(4, 3)
        self.voxel_dim = voxel_dim

    def texture(self, which):
        """(data, offsets, widths, heights, strides, type) of 'cubemap' | 'cubemap_specular' | 'brdf_lut'."""
        t = getattr(self.env, which).contents
        nm = t.max_mip_level + 1
        arr = lambda p, n: np.ctypeslib.as_array(p, (n,)).copy()
        return arr(t.data, t.ntexels), arr(t.mip_offsets, nm), arr(t.mip_widths, nm), arr(t.mip_heights, nm), arr(t.array_stride, nm), t.texture_type

    def voxels(self):
        n = self.voxel_dim ** 3
        return np.ctypeslib.as_array(self.env.voxel_grid.gi_sh4, (n * 16,)).copy().reshape(n, 4, 4)

    def close(self):
        if self._h:
            self._host.swrh_env_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ggx_cache_load(path, width, height):
    """The reference's `.ggx` cache (texture.rs:426-514): (mips, 6, height, width) uint32 texels, or None when the file is missing
    or is a cache of something else (the reference then bakes and writes one)."""
    host = _host()
    host.swrh_ggx_cache_load.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
    mips = 1 + int(max(width, height)).bit_length() - 1
    out = np.zeros(width * height * 6 * mips, np.uint32)
    rc = host.swrh_ggx_cache_load(os.fsencode(path), width, height, out.ctypes.data)
    if rc < 0:
        raise GltfError(host.swrh_last_error().decode())
    return out.reshape(mips, 6, height, width) if rc == 1 else None


def ggx_cache_save(path, texels):
    """texels: (mips, 6, height, width) uint32 (texture.rs:516-552)."""
    host = _host()
    host.swrh_ggx_cache_save.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    t = np.ascontiguousarray(texels, np.uint32)
    assert t.ndim == 4 and t.shape[1] == 6
    _check(host.swrh_ggx_cache_save(os.fsencode(path), t.shape[3], t.shape[2], t.shape[0], t.ctypes.data), host)


def gi_cache_load(path, dims):
    """The reference's `.gi` cache (gi.rs:30-83) for a (w, h, d) voxel grid: (w*h*d, 4, 4) float32 SH4 coefficients (r, g, b, w), or
    None when there is no usable cache."""
    host = _host()
    host.swrh_gi_cache_load.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    w, h, d = (int(x) for x in dims)
    out = np.zeros(w * h * d * 16, np.float32)
    rc = host.swrh_gi_cache_load(os.fsencode(path), w, h, d, out.ctypes.data)
    if rc < 0:
        raise GltfError(host.swrh_last_error().decode())
    return out.reshape(w * h * d, 4, 4) if rc == 1 else None


def gi_cache_save(path, dims, gi_sh4):
    """gi.rs:85-118."""
    host = _host()
    host.swrh_gi_cache_save.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    w, h, d = (int(x) for x in dims)
    g = np.ascontiguousarray(gi_sh4, np.float32)
    assert g.size == w * h * d * 16
    _check(host.swrh_gi_cache_save(os.fsencode(path), w, h, d, g.ctypes.data), host)


def load_scene(path, sky_cross_rgba, grid_size=128, lut_size=128, specular_samples=64, ggx_cache=None):
    """What the viewer's load_scene does before the first frame (main.rs:100-291), for the fields the renderer consumes:
    parse the glTF / GLB file, bake the environment from the sky image (scene.rs:151-231), span a grid_size^3 voxel grid over
    the scene bounds (main.rs:228-235), initialise it from the irradiance SH with GI_FALLBACK_SCENE_SH_SCALE = 0.25 and sky
    visibility 1 (main.rs:54, :280; gi.rs:123-149) and fill in the ray-cast sun visibility (main.rs:237-246).
    Returns (scene, default camera spec as main.rs:210-224 builds it: (position, look_at, fov, far_plane))."""
    env = BakedEnvironment(sky_cross_rgba, lut_size=lut_size, specular_samples=specular_samples, voxel_dim=grid_size, irradiance_scale=0.25,
                           sky_visibility=1.0, light_intensity=1.0, ggx_cache=ggx_cache)
    scene = load_gltf(path, environment=env)
    bake_sun_visibility(scene)
    if scene.cameras:
        # main.rs:198-209 + RenderCamera::from_gltf (rendercamera.rs:65-76): only the FIRST camera's yfov is used; position,
        # target and far plane are fixed (the node transform is ignored); an orthographic camera is a load error
        first = scene.cameras[0]
        if not first["perspective"]:
            raise GltfError("Failed to create camera from GLTF: unsupported camera type")
        return scene, ((0.0, 0.0, 5.0), (0.0, 0.0, 0.0), float(first["yfov_or_xmag"]), 1000.0)
    c, d = scene.bounds_center, float(scene.bounds_diagonal)
    camera = ((float(c[0]), float(c[1]), float(c[2]) + d), (float(c[0]), float(c[1]), float(c[2])), float(np.float32(np.pi) / np.float32(4.0)), d * 2.0)
    return scene, camera


def compute_sun_visibility(scene):
    """Blurred sun visibility per voxel of `scene` (anything with .desc()), shape (d, h, w): gi.rs:151-314 on the host."""
    host = _host()
    host.swrh_compute_sun_visibility.argtypes = [C.POINTER(abi.SceneDesc), C.c_void_p]
    d = scene.desc()
    w, h, dd = d.voxel_grid.dims[:]
    out = np.zeros(w * h * dd, np.float32)
    _check(host.swrh_compute_sun_visibility(C.byref(d), out.ctypes.data), host)
    return out.reshape(dd, h, w)


def bake_sun_visibility(gltf_scene):
    """Write the sun visibility into a loaded document's voxel grid (gi_sh4[v][0].w) in place, before its first render."""
    host = _host()
    host.swrh_gltf_bake_sun_visibility.argtypes = [C.c_void_p]
    _check(host.swrh_gltf_bake_sun_visibility(gltf_scene._h), host)

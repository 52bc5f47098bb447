"""Python face of the host mirror (host/swr_host.hpp), same names as the reference:
Renderer::new / render_scene / update_auto_exposure / blit_to_buffer (renderer.rs:165-355),
RenderCamera::new (rendercamera.rs:28), RenderBuffer (renderer.rs:24-28).

All compute happens in libswr_b200.so (CUDA). There is no fallback: if the libraries are
missing or no device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("SWR_LIB_DIR") or os.path.join(_HERE, "lib")  # SWR_LIB_DIR: A/B builds (tools/build_variant.sh)


class LibraryMissing(RuntimeError):
    pass


_libs = None


def load_libraries():
    """Load libswr_b200.so (C ABI, CUDA) and libswr_host.so (C++ host mirror). Raises if absent."""
    global _libs
    if _libs is not None:
        return _libs
    core_p = os.path.join(LIB_DIR, "libswr_b200.so")
    host_p = os.path.join(LIB_DIR, "libswr_host.so")
    for p in (core_p, host_p):
        if not os.path.exists(p):
            raise LibraryMissing(f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback for the rasterisation path)")
    core = C.CDLL(core_p, mode=C.RTLD_GLOBAL)
    host = C.CDLL(host_p, mode=C.RTLD_GLOBAL)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    core.swr_abi_version.restype = i32
    core.swr_last_error.restype = C.c_char_p
    core.swr_last_error.argtypes = [vp]
    core.swr_create.restype = vp
    core.swr_create.argtypes = [i32, i32, i32]
    core.swr_destroy.argtypes = [vp]
    core.swr_set_tile_rows.argtypes = [vp, i32, i32]
    core.swr_set_rsqrt_table.argtypes = [vp, vp, i32]
    core.swr_upload_scene.argtypes = [vp, C.POINTER(abi.SceneDesc)]
    core.swr_render.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.Draw), i32, i32]
    core.swr_set_fixed_exposure.argtypes = [vp, f32]
    core.swr_shade.argtypes = [vp, C.POINTER(abi.Camera)]
    core.swr_keys_to_global.argtypes = [vp]
    core.swr_keys_localize.argtypes = [vp]
    core.swr_shade_composited.argtypes = [vp, C.POINTER(abi.Camera), i32, i32]
    core.swr_device_bary.restype = vp
    core.swr_device_bary.argtypes = [vp]
    core.swr_resolve.argtypes = [vp, f32, vp]
    core.swr_read_tile_luminance.argtypes = [vp, vp]
    core.swr_read_tile_costs.argtypes = [vp, vp, vp]
    core.swr_read_visbuffer.argtypes = [vp, vp, vp, vp, vp]
    core.swr_read_color.argtypes = [vp, vp]
    core.swr_synchronize.argtypes = [vp]
    core.swr_get_stats.argtypes = [vp, C.POINTER(abi.FrameStats)]
    for n in ("swr_device_pixels", "swr_device_keys", "swr_cuda_stream"):
        getattr(core, n).restype = vp
        getattr(core, n).argtypes = [vp]
    core.swr_device_keys_bytes.restype = C.c_size_t
    core.swr_device_keys_bytes.argtypes = [vp]
    core.swr_peer_export.argtypes = [vp, vp]
    core.swr_peer_open.argtypes = [vp, vp]
    core.swr_peer_attach.argtypes = [vp, vp]
    core.swr_resolve_peer.argtypes = [vp, f32, C.c_uint32]
    core.swr_peer_collect.argtypes = [vp, C.c_uint32, i32]
    core.swr_peer_release.argtypes = [vp, C.c_uint32]
    core.swr_launch_count.restype = C.c_uint64
    core.swr_launch_count.argtypes = [vp]
    core.swr_sizeof.restype = C.c_size_t
    core.swr_sizeof.argtypes = [i32]

    host.swrh_last_error.restype = C.c_char_p
    host.swrh_camera_build.argtypes = [C.POINTER(f32), C.POINTER(f32), f32, f32, f32, f32, C.POINTER(abi.Camera)]
    host.swrh_camera_build_rotated.argtypes = [C.POINTER(f32), C.POINTER(f32), f32, f32, f32, f32, f32, f32, C.POINTER(abi.Camera)]
    host.swrh_renderer_new.restype = vp
    host.swrh_renderer_new.argtypes = [i32, i32, i32]
    host.swrh_renderer_new_lanes.restype = vp
    host.swrh_renderer_new_lanes.argtypes = [i32, i32, i32, i32]
    host.swrh_renderer_lanes.argtypes = [vp]
    host.swrh_renderer_lane_ctx.restype = vp
    host.swrh_renderer_lane_ctx.argtypes = [vp, i32]
    host.swrh_renderer_new_multi.restype = vp
    host.swrh_renderer_new_multi.argtypes = [i32, i32, C.POINTER(i32), i32]
    host.swrh_renderer_tile_rows.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    host.swrh_renderer_free.argtypes = [vp]
    host.swrh_renderer_ctx.restype = vp
    host.swrh_renderer_ctx.argtypes = [vp]
    host.swrh_set_tile_rows.argtypes = [vp, i32, i32]
    host.swrh_invalidate_scene.argtypes = [vp]
    host.swrh_set_reference_rsqrt.argtypes = [vp, i32]
    host.swrh_reference_rsqrt_bits.argtypes = [vp]
    host.swrh_render_scene.argtypes = [vp, C.POINTER(abi.SceneDesc), C.POINTER(abi.Camera), i32, i32, i32]
    host.swrh_update_auto_exposure.argtypes = [vp, f32]
    host.swrh_frame_exposure.argtypes = [vp, f32]
    host.swrh_frame_hdr.argtypes = [vp]
    host.swrh_num_draws.argtypes = [vp]
    host.swrh_auto_exposure_step.argtypes = [vp, vp, i32, f32]
    host.swrh_auto_exposure.restype = f32
    host.swrh_auto_exposure.argtypes = [vp]
    host.swrh_blit_to_buffer.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    host.swrh_blit_to_buffer_async.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.POINTER(i32)]
    host.swrh_wait_blit.argtypes = [vp, i32]
    host.swrh_build_draws.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(abi.Camera), C.POINTER(abi.Draw), i32, i32, i32]
    _libs = (core, host)
    return _libs


class RenderCamera:
    """rendercamera.rs:28-63 `RenderCamera::new(position, look_at, fov, width, height, far_plane)`."""

    def __init__(self, position, look_at, fov, width, height, far_plane):
        _, host = load_libraries()
        self.abi = abi.Camera()
        pos = (C.c_float * 3)(*position)
        la = (C.c_float * 3)(*look_at)
        if host.swrh_camera_build(pos, la, fov, float(width), float(height), far_plane, C.byref(self.abi)) != 0:
            raise RuntimeError(host.swrh_last_error().decode())
        self.width, self.height = width, height

    @classmethod
    def from_spec(cls, spec, width, height):
        """Aim at spec.look_at the way the reference's controls do: RenderCamera::new towards a level
        target (yaw only), then rotate_mouse for the pitch (rendercamera.rs:119-123)."""
        import math
        _, host = load_libraries()
        f = [spec.look_at[i] - spec.position[i] for i in range(3)]
        n = math.sqrt(sum(c * c for c in f))
        f = [c / n for c in f]
        self = cls.__new__(cls)
        self.abi = abi.Camera()
        self.width, self.height = width, height
        level = (spec.position[0] + f[0], spec.position[1], spec.position[2] + f[2])
        if abs(f[0]) + abs(f[2]) < 1e-6:
            level = (spec.position[0], spec.position[1], spec.position[2] - 1.0)
        mouse_dy = -math.asin(max(-1.0, min(1.0, f[1]))) / 0.01
        pos = (C.c_float * 3)(*spec.position)
        la = (C.c_float * 3)(*level)
        if host.swrh_camera_build_rotated(pos, la, 0.0, mouse_dy, spec.fov, float(width), float(height), spec.far_plane, C.byref(self.abi)) != 0:
            raise RuntimeError(host.swrh_last_error().decode())
        return self


class RenderBuffer:
    """renderer.rs:24-28: caller-owned W*H u32 pixels, row-major, (R<<24)|(G<<16)|(B<<8)|A."""

    def __init__(self, width, height, pixels=None, pinned=False):
        self.width, self.height = width, height
        if pixels is None:
            if pinned:
                import torch
                self._t = torch.empty(width * height, dtype=torch.int32).pin_memory()
                pixels = self._t.numpy().view(np.uint32)
            else:
                pixels = np.zeros(width * height, np.uint32)
        self.pixels = pixels

    def clear(self):
        self.pixels.fill(0)

    def set_pixel(self, x, y, color):
        if x < self.width and y < self.height:
            self.pixels[y * self.width + x] = color


class Renderer:
    """renderer.rs:145-355 behind the CUDA path. One Renderer per GPU."""

    def __init__(self, width, height, device=0, devices=None, lanes=1):
        """device: one CUDA ordinal; devices=[...]: one Renderer over several GPUs of this process (sort-first);
        lanes=2: frames alternate between two contexts of the device (for pipelined callers: blit_to_buffer_async)."""
        self.core, self.host = load_libraries()
        self.width, self.height = width, height
        self.tiles_x = (width + abi.TILE_SIZE - 1) // abi.TILE_SIZE
        self.tiles_y = (height + abi.TILE_SIZE - 1) // abi.TILE_SIZE
        self.devices = list(devices) if devices is not None else None
        if self.devices is not None:
            arr = (C.c_int * len(self.devices))(*self.devices)
            self._h = self.host.swrh_renderer_new_multi(width, height, arr, len(self.devices))
        elif lanes != 1:
            self._h = self.host.swrh_renderer_new_lanes(width, height, device, lanes)
        else:
            self._h = self.host.swrh_renderer_new(width, height, device)
        if not self._h:
            raise RuntimeError("Renderer::new failed: " + self.host.swrh_last_error().decode())
        self._scene = None

    @property
    def ctx(self):
        """The device context of the frame rendered last (the lanes of a pipelined renderer alternate)."""
        return self.host.swrh_renderer_ctx(self._h)

    def close(self):
        if self._h:
            self.host.swrh_renderer_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.host.swrh_last_error().decode())

    def _check_core(self, rc):
        if rc != 0:
            raise RuntimeError(self.core.swr_last_error(self.ctx).decode())

    def set_reference_rsqrt(self, on):
        """normalize() as the reference computes it on this host (_mm_rsqrt_ps table; default) or rsqrtf()."""
        self._check(self.host.swrh_set_reference_rsqrt(self._h, int(on)))

    @property
    def reference_rsqrt_bits(self):
        return self.host.swrh_reference_rsqrt_bits(self._h)

    def set_tile_rows(self, r0, r1):
        self._check(self.host.swrh_set_tile_rows(self._h, r0, r1))

    def render_scene(self, scene, camera, shade=True, shard=0, nshards=1):
        """scene: scenes.SceneData; camera: RenderCamera."""
        desc = scene.desc()
        if scene is not self._scene:
            # a new SceneData may be handed a descriptor at a recycled address: never trust pointer identity across scenes
            self._check(self.host.swrh_invalidate_scene(self._h))
        self._check(self.host.swrh_render_scene(self._h, C.byref(desc), C.byref(camera.abi), int(shade), shard, nshards))
        self._scene = scene  # keep the arrays alive while the context references the descriptor

    @property
    def num_draws(self):
        """Draws submitted by the last render_scene (after frustum / band culling)."""
        return self.host.swrh_num_draws(self._h)

    def update_auto_exposure(self, delta_time):
        self._check(self.host.swrh_update_auto_exposure(self._h, delta_time))

    @property
    def auto_exposure(self):
        return self.host.swrh_auto_exposure(self._h)

    def blit_to_buffer(self, buffer):
        self._check(self.host.swrh_blit_to_buffer(self._h, buffer.pixels.ctypes.data, buffer.width, buffer.height))

    def blit_to_buffer_async(self, buffer):
        """Start resolve + read-back into `buffer` (pinned) without blocking; returns a ticket for wait_blit."""
        t = C.c_int(-1)
        self._check(self.host.swrh_blit_to_buffer_async(self._h, buffer.pixels.ctypes.data, buffer.width, buffer.height, C.byref(t)))
        return t.value

    def wait_blit(self, ticket):
        self._check(self.host.swrh_wait_blit(self._h, ticket))

    # ---- C-ABI extras used by tests / bench -----------------------------------------------
    def resolve_device_only(self, exposure=abi.DEFAULT_EXPOSURE):
        self._check(self.host.swrh_frame_exposure(self._h, exposure))
        self._check_core(self.core.swr_resolve(self.ctx, exposure, None))

    def shade(self, camera):
        self._check_core(self.core.swr_shade(self.ctx, C.byref(camera.abi)))

    # ---- sort-last steps (see include/swr.h) ----------------------------------------------------
    def keys_to_global(self):
        self._check_core(self.core.swr_keys_to_global(self.ctx))

    def keys_localize(self):
        self._check_core(self.core.swr_keys_localize(self.ctx))

    def shade_composited(self, camera, sky_row_begin, sky_row_end):
        self._check_core(self.core.swr_shade_composited(self.ctx, C.byref(camera.abi), sky_row_begin, sky_row_end))

    # ---- sort-first frame assembly over peer memory (see include/swr.h) ---------------------------
    def peer_export(self):
        """Assembling rank: bytes of the handle the contributing ranks open."""
        h = (C.c_ubyte * abi.PEER_HANDLE_BYTES)()
        self._check_core(self.core.swr_peer_export(self.ctx, h))
        return bytes(h)

    def peer_open(self, handle):
        h = (C.c_ubyte * abi.PEER_HANDLE_BYTES).from_buffer_copy(handle)
        self._check_core(self.core.swr_peer_open(self.ctx, h))

    def peer_attach(self, device_ptr):
        self._check_core(self.core.swr_peer_attach(self.ctx, device_ptr))

    def resolve_peer(self, exposure, frame):
        self._check(self.host.swrh_frame_exposure(self._h, exposure))
        self._check_core(self.core.swr_resolve_peer(self.ctx, exposure, frame))

    def peer_collect(self, frame, contributors):
        self._check_core(self.core.swr_peer_collect(self.ctx, frame, contributors))

    def peer_release(self, frame):
        self._check_core(self.core.swr_peer_release(self.ctx, frame))

    def device_bary_ptr(self):
        return self.core.swr_device_bary(self.ctx)

    def synchronize(self):
        self._check_core(self.core.swr_synchronize(self.ctx))

    def read_visbuffer(self):
        n = self.width * self.height
        depth, seq = np.empty(n, np.uint32), np.empty(n, np.uint32)
        b1, b2 = np.empty(n, np.float32), np.empty(n, np.float32)
        self._check_core(self.core.swr_read_visbuffer(self.ctx, depth.ctypes.data, seq.ctypes.data, b1.ctypes.data, b2.ctypes.data))
        return depth, seq, b1, b2

    def read_color(self):
        rgb = np.empty(self.width * self.height * 3, np.float32)
        self._check(self.host.swrh_frame_hdr(self._h))
        self._check_core(self.core.swr_read_color(self.ctx, rgb.ctypes.data))
        return rgb.reshape(self.height, self.width, 3)

    def read_tile_luminance(self):
        out = np.empty(self.tiles_x * self.tiles_y, np.float32)
        self._check_core(self.core.swr_read_tile_luminance(self.ctx, out.ctypes.data))
        return out

    def read_tile_costs(self):
        """(refs, raster cycles) per tile of the last frame, each (tiles_y, tiles_x)."""
        refs = np.empty(self.tiles_x * self.tiles_y, np.uint32)
        cyc = np.empty(self.tiles_x * self.tiles_y, np.uint32)
        self._check_core(self.core.swr_read_tile_costs(self.ctx, refs.ctypes.data, cyc.ctypes.data))
        return refs.reshape(self.tiles_y, self.tiles_x), cyc.reshape(self.tiles_y, self.tiles_x)

    def read_tile_counts(self):
        return self.read_tile_costs()[0]

    def stats(self):
        st = abi.FrameStats()
        self._check_core(self.core.swr_get_stats(self.ctx, C.byref(st)))
        return st.as_dict()

    def device_tile_rows(self, i):
        """Rows [begin, end) device i of a multi-device renderer owns."""
        a, b = C.c_int(), C.c_int()
        self._check(self.host.swrh_renderer_tile_rows(self._h, i, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def launch_count(self):
        """CUDA kernels launched by this renderer's context(s) so far (swr_launch_count, summed over the lanes)."""
        n = self.host.swrh_renderer_lanes(self._h)
        if n <= 0:
            return int(self.core.swr_launch_count(self.ctx))
        return sum(int(self.core.swr_launch_count(self.host.swrh_renderer_lane_ctx(self._h, i))) for i in range(n))

    def device_pixels_ptr(self):
        return self.core.swr_device_pixels(self.ctx)

    def device_keys_ptr(self):
        return self.core.swr_device_keys(self.ctx), self.core.swr_device_keys_bytes(self.ctx)

    def cuda_stream(self):
        return self.core.swr_cuda_stream(self.ctx)


def build_draws(scene, camera, shard=0, nshards=1, max_draws=1 << 20):
    """Host draw list (renderer.rs:357-468) without touching a device."""
    _, host = load_libraries()
    n = host.swrh_build_draws(C.byref(scene.desc()), C.byref(camera.abi), None, 0, shard, nshards)
    if n < 0:
        raise RuntimeError(host.swrh_last_error().decode())
    arr = (abi.Draw * max(n, 1))()
    host.swrh_build_draws(C.byref(scene.desc()), C.byref(camera.abi), arr, n, shard, nshards)
    return arr, n

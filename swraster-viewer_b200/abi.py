"""ctypes mirror of include/swr.h (the C ABI of libswr_b200.so).

Only plain C structs live here; no compute. Field order and types must match
include/swr.h exactly — tests/test_abi.py checks sizeof() against the library.
"""
import ctypes as C

TILE_SIZE = 64
DEFAULT_EXPOSURE = 2.0

TEX_SRGB, TEX_NORMAL, TEX_METALLIC_ROUGHNESS, TEX_CUBEMAP, TEX_LINEAR = range(5)
WRAP_REPEAT, WRAP_MIRRORED_REPEAT, WRAP_CLAMP_TO_EDGE = range(3)
MAT_ALPHA_TESTED, MAT_TRANSLUCENT = 1, 2
DRAW_CLIP = 1

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)


class PrimitiveDesc(C.Structure):
    _fields_ = [
        ("positions", f32p), ("normals", f32p), ("tangents", f32p), ("texcoords", f32p),
        ("indices", u32p), ("nverts", C.c_uint32), ("nindices", C.c_uint32),
        ("material_index", C.c_uint32), ("bounding_sphere", C.c_float * 4),
    ]


class MeshDesc(C.Structure):
    _fields_ = [("first_primitive", C.c_uint32), ("num_primitives", C.c_uint32)]


class NodeDesc(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("mesh_index", C.c_int32),
                ("bounding_sphere_world", C.c_float * 4)]


class TextureDesc(C.Structure):
    _fields_ = [
        ("data", u32p), ("ntexels", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
        ("texture_type", C.c_uint32), ("max_mip_level", C.c_uint32),
        ("mip_offsets", u32p), ("mip_widths", u32p), ("mip_heights", u32p), ("array_stride", u32p),
        ("wrap_s", C.c_uint32), ("wrap_t", C.c_uint32),
    ]


class MaterialDesc(C.Structure):
    _fields_ = [
        ("base_color_factor", C.c_float * 4), ("metallic_factor", C.c_float), ("roughness_factor", C.c_float),
        ("emissive_factor", C.c_float * 3), ("occlusion_strength", C.c_float), ("transmission", C.c_float),
        ("alpha_cutoff", C.c_float), ("flags", C.c_uint32),
        ("base_color_texture", C.c_int32), ("metallic_roughness_texture", C.c_int32),
        ("normal_texture", C.c_int32), ("emissive_texture", C.c_int32),
        ("occlusion_texture", C.c_int32), ("transmission_texture", C.c_int32),
    ]


class VoxelGridDesc(C.Structure):
    _fields_ = [("dims", C.c_uint32 * 3), ("world_min", C.c_float * 3), ("world_max", C.c_float * 3),
                ("gi_sh4", f32p)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("primitives", C.POINTER(PrimitiveDesc)), ("nprimitives", C.c_uint32),
        ("meshes", C.POINTER(MeshDesc)), ("nmeshes", C.c_uint32),
        ("nodes", C.POINTER(NodeDesc)), ("nnodes", C.c_uint32),
        ("materials", C.POINTER(MaterialDesc)), ("nmaterials", C.c_uint32),
        ("textures", C.POINTER(TextureDesc)), ("ntextures", C.c_uint32),
        ("voxel_grid", VoxelGridDesc),
        ("cubemap", C.c_int32), ("cubemap_specular", C.c_int32), ("brdf_lut", C.c_int32),
        ("light_direction", C.c_float * 3), ("light_color", C.c_float * 3),
    ]


class Camera(C.Structure):
    _fields_ = [
        ("position", C.c_float * 4), ("view_matrix", C.c_float * 16), ("view_project_matrix", C.c_float * 16),
        ("skybox_matrix_transposed", C.c_float * 16), ("view_clip_planes", (C.c_float * 4) * 6),
        ("one_over_width", C.c_float), ("one_over_height", C.c_float), ("reserved", C.c_float * 2),
    ]


class Draw(C.Structure):
    _fields_ = [("model", C.c_float * 16), ("mvp", C.c_float * 16), ("primitive", C.c_uint32),
                ("flags", C.c_uint32), ("first_triangle", C.c_uint32), ("reserved", C.c_uint32)]


class FrameStats(C.Structure):
    _fields_ = [
        ("triangles_submitted", C.c_uint64), ("vertices_submitted", C.c_uint64),
        ("triangles_binned", C.c_uint64), ("triangles_clipped", C.c_uint64), ("tile_refs", C.c_uint64),
        ("tiles", C.c_uint32), ("clusters_culled", C.c_uint32),
        ("ms_setup_bin", C.c_float), ("ms_raster", C.c_float), ("ms_shade", C.c_float), ("ms_resolve", C.c_float),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


# Every symbol include/swr.h declares (tests check the library exports all of them).
PEER_HANDLE_BYTES = 64  # SWR_PEER_HANDLE_BYTES

EXPORTS = [
    "swr_abi_version", "swr_last_error", "swr_create", "swr_destroy", "swr_set_tile_rows", "swr_set_rsqrt_table", "swr_upload_scene", "swr_share_scene",
    "swr_render", "swr_set_fixed_exposure", "swr_shade", "swr_keys_to_global", "swr_keys_localize", "swr_shade_composited", "swr_device_bary", "swr_resolve", "swr_resolve_async", "swr_wait_pixels", "swr_read_tile_luminance", "swr_read_tile_costs", "swr_read_visbuffer", "swr_read_color",
    "swr_synchronize", "swr_get_stats", "swr_device_pixels", "swr_device_keys", "swr_device_keys_bytes",
    "swr_cuda_stream", "swr_sizeof", "swr_launch_count",
    "swr_peer_export", "swr_peer_open", "swr_peer_attach", "swr_resolve_peer", "swr_peer_collect", "swr_peer_release",
    "swr_multi_create", "swr_multi_destroy", "swr_multi_last_error", "swr_multi_device_count", "swr_multi_context", "swr_multi_tile_rows",
    "swr_multi_set_rsqrt_table", "swr_multi_upload_scene", "swr_multi_render", "swr_multi_resolve", "swr_multi_read_tile_luminance",
    "swr_multi_get_stats", "swr_multi_synchronize",
    "swr_bake_brdf_lut", "swr_bake_irradiance_sh4", "swr_bake_prefilter_specular", "swr_bake_sun_visibility", "swr_bake_last_error",
]

//! `extern "C"` declarations for libswr_b200.so — one Rust item per declaration of include/swr.h, same order.
//!
//! This image has no Rust toolchain (cargo / rustc absent, crates not vendored), so this file cannot be compiled here.
//! What IS checked here: tests/test_rust_shim.py parses the `extern "C"` block below and include/swr.h and fails when a
//! function name, its argument count or an argument's pointer-ness differs, and when a `#[repr(C)]` struct's field list
//! differs from the header's. The C++ mirror (swraster-viewer_b200/host/swr_host.hpp) exercises the same calls on the GPU.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const SWR_OK: c_int = 0;
pub const SWR_MAT_ALPHA_TESTED: u32 = 1;
pub const SWR_MAT_TRANSLUCENT: u32 = 2;
pub const SWR_DRAW_CLIP: u32 = 1;
pub const SWR_DRAW_TRANSLUCENT: u32 = 2;
pub const SWR_MULTI_SORT_FIRST: c_int = 1;
pub const SWR_PEER_HANDLE_BYTES: usize = 64;
pub const SWR_TEX_SRGB: u32 = 0;
pub const SWR_TEX_NORMAL: u32 = 1;
pub const SWR_TEX_METALLIC_ROUGHNESS: u32 = 2;
pub const SWR_TEX_CUBEMAP: u32 = 3;
pub const SWR_TEX_LINEAR: u32 = 4;
pub const SWR_WRAP_REPEAT: u32 = 0;
pub const SWR_WRAP_MIRRORED_REPEAT: u32 = 1;
pub const SWR_WRAP_CLAMP_TO_EDGE: u32 = 2;

#[repr(C)]
pub struct swr_primitive_desc {
    pub positions: *const f32, // glam Vec4 per vertex (scene.rs:95)
    pub normals: *const f32,   // glam Vec3A per vertex, 16-byte stride (scene.rs:96)
    pub tangents: *const f32,  // glam Vec4 (scene.rs:97)
    pub texcoords: *const f32, // glam Vec2 (scene.rs:98)
    pub indices: *const u32,
    pub nverts: u32,
    pub nindices: u32,
    pub material_index: u32,
    pub bounding_sphere: [f32; 4],
}
#[repr(C)]
pub struct swr_mesh_desc {
    pub first_primitive: u32,
    pub num_primitives: u32,
}
#[repr(C)]
pub struct swr_node_desc {
    pub transform: [f32; 16],
    pub mesh_index: i32,
    pub bounding_sphere_world: [f32; 4],
}
#[repr(C)]
pub struct swr_texture_desc {
    pub data: *const u32,
    pub ntexels: u32,
    pub width: u32,
    pub height: u32,
    pub texture_type: u32,
    pub max_mip_level: u32,
    pub mip_offsets: *const u32,
    pub mip_widths: *const u32,
    pub mip_heights: *const u32,
    pub array_stride: *const u32,
    pub wrap_s: u32,
    pub wrap_t: u32,
}
#[repr(C)]
pub struct swr_material_desc {
    pub base_color_factor: [f32; 4],
    pub metallic_factor: f32,
    pub roughness_factor: f32,
    pub emissive_factor: [f32; 3],
    pub occlusion_strength: f32,
    pub transmission: f32,
    pub alpha_cutoff: f32,
    pub flags: u32,
    pub base_color_texture: i32,
    pub metallic_roughness_texture: i32,
    pub normal_texture: i32,
    pub emissive_texture: i32,
    pub occlusion_texture: i32,
    pub transmission_texture: i32,
}
#[repr(C)]
pub struct swr_voxel_grid_desc {
    pub dims: [u32; 3],
    pub world_min: [f32; 3],
    pub world_max: [f32; 3],
    pub gi_sh4: *const f32,
}
#[repr(C)]
pub struct swr_scene_desc {
    pub primitives: *const swr_primitive_desc,
    pub nprimitives: u32,
    pub meshes: *const swr_mesh_desc,
    pub nmeshes: u32,
    pub nodes: *const swr_node_desc,
    pub nnodes: u32,
    pub materials: *const swr_material_desc,
    pub nmaterials: u32,
    pub textures: *const swr_texture_desc,
    pub ntextures: u32,
    pub voxel_grid: swr_voxel_grid_desc,
    pub cubemap: i32,
    pub cubemap_specular: i32,
    pub brdf_lut: i32,
    pub light_direction: [f32; 3],
    pub light_color: [f32; 3],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct swr_camera {
    pub position: [f32; 4],
    pub view_matrix: [f32; 16],
    pub view_project_matrix: [f32; 16],
    pub skybox_matrix_transposed: [f32; 16],
    pub view_clip_planes: [[f32; 4]; 6],
    pub one_over_width: f32,
    pub one_over_height: f32,
    pub reserved: [f32; 2],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct swr_draw {
    pub model: [f32; 16],
    pub mvp: [f32; 16],
    pub primitive: u32,
    pub flags: u32,
    pub first_triangle: u32,
    pub reserved: u32,
}
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct swr_frame_stats {
    pub triangles_submitted: u64,
    pub vertices_submitted: u64,
    pub triangles_binned: u64,
    pub triangles_clipped: u64,
    pub tile_refs: u64,
    pub tiles: u32,
    pub clusters_culled: u32,
    pub ms_setup_bin: f32,
    pub ms_raster: f32,
    pub ms_shade: f32,
    pub ms_resolve: f32,
}
#[repr(C)]
pub struct swr_bvh_node {
    pub lo: [f32; 3],
    pub hi: [f32; 3],
    pub first: u32,
    pub count: u32,
}
#[repr(C)]
pub struct swr_sun_triangle {
    pub p0: [f32; 3],
    pub p1: [f32; 3],
    pub p2: [f32; 3],
    pub transmission: f32,
}
#[repr(C)]
pub struct swr_sunvis_desc {
    pub nodes: *const swr_bvh_node,
    pub nnodes: u32,
    pub order: *const u32,
    pub norder: u32,
    pub triangles: *const swr_sun_triangle,
    pub ntriangles: u32,
    pub active: *const u8,
    pub dims: [u32; 3],
    pub world_min: [f32; 3],
    pub world_max: [f32; 3],
    pub light_direction: [f32; 3],
}
#[repr(C)]
pub struct swr_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct swr_multi {
    _private: [u8; 0],
}

#[link(name = "swr_b200")]
extern "C" {
    pub fn swr_abi_version() -> c_int;
    pub fn swr_sizeof(which: c_int) -> usize;
    pub fn swr_last_error(ctx: *const swr_ctx) -> *const c_char;
    pub fn swr_create(width: c_int, height: c_int, device: c_int) -> *mut swr_ctx;
    pub fn swr_destroy(ctx: *mut swr_ctx);
    pub fn swr_set_tile_rows(ctx: *mut swr_ctx, row_begin: c_int, row_end: c_int) -> c_int;
    pub fn swr_set_rsqrt_table(ctx: *mut swr_ctx, table: *const u32, mantissa_bits: c_int) -> c_int;
    pub fn swr_upload_scene(ctx: *mut swr_ctx, scene: *const swr_scene_desc) -> c_int;
    pub fn swr_share_scene(ctx: *mut swr_ctx, owner: *const swr_ctx) -> c_int;
    pub fn swr_render(ctx: *mut swr_ctx, camera: *const swr_camera, draws: *const swr_draw, ndraws: c_int, shade: c_int) -> c_int;
    pub fn swr_set_fixed_exposure(ctx: *mut swr_ctx, exposure: c_float) -> c_int;
    pub fn swr_shade(ctx: *mut swr_ctx, camera: *const swr_camera) -> c_int;
    pub fn swr_keys_to_global(ctx: *mut swr_ctx) -> c_int;
    pub fn swr_keys_localize(ctx: *mut swr_ctx) -> c_int;
    pub fn swr_shade_composited(ctx: *mut swr_ctx, camera: *const swr_camera, sky_row_begin: c_int, sky_row_end: c_int) -> c_int;
    pub fn swr_device_bary(ctx: *mut swr_ctx) -> *mut c_void;
    pub fn swr_resolve(ctx: *mut swr_ctx, exposure: f32, out_pixels: *mut u32) -> c_int;
    pub fn swr_resolve_async(ctx: *mut swr_ctx, exposure: f32, out_pixels: *mut u32, ticket: *mut c_int) -> c_int;
    pub fn swr_wait_pixels(ctx: *mut swr_ctx, ticket: c_int) -> c_int;
    pub fn swr_read_tile_costs(ctx: *mut swr_ctx, refs_per_tile: *mut u32, raster_cycles_per_tile: *mut u32) -> c_int;
    pub fn swr_read_tile_luminance(ctx: *mut swr_ctx, out_per_tile: *mut f32) -> c_int;
    pub fn swr_read_visbuffer(ctx: *mut swr_ctx, depth_bits: *mut u32, seq: *mut u32, bary1: *mut f32, bary2: *mut f32) -> c_int;
    pub fn swr_read_color(ctx: *mut swr_ctx, rgb: *mut f32) -> c_int;
    pub fn swr_synchronize(ctx: *mut swr_ctx) -> c_int;
    pub fn swr_get_stats(ctx: *mut swr_ctx, out: *mut swr_frame_stats) -> c_int;
    pub fn swr_launch_count(ctx: *mut swr_ctx) -> u64;
    pub fn swr_peer_export(ctx: *mut swr_ctx, handle_out: *mut c_void) -> c_int;
    pub fn swr_peer_open(ctx: *mut swr_ctx, handle: *const c_void) -> c_int;
    pub fn swr_peer_attach(ctx: *mut swr_ctx, assembler_device_pixels: *mut c_void) -> c_int;
    pub fn swr_resolve_peer(ctx: *mut swr_ctx, exposure: f32, frame: u32) -> c_int;
    pub fn swr_peer_collect(ctx: *mut swr_ctx, frame: u32, contributors: c_int) -> c_int;
    pub fn swr_peer_release(ctx: *mut swr_ctx, frame: u32) -> c_int;
    pub fn swr_multi_create(width: c_int, height: c_int, devices: *const c_int, ndev: c_int, mode: c_int) -> *mut swr_multi;
    pub fn swr_multi_destroy(m: *mut swr_multi);
    pub fn swr_multi_last_error(m: *const swr_multi) -> *const c_char;
    pub fn swr_multi_device_count(m: *const swr_multi) -> c_int;
    pub fn swr_multi_context(m: *mut swr_multi, i: c_int) -> *mut swr_ctx;
    pub fn swr_multi_tile_rows(m: *const swr_multi, i: c_int, row_begin: *mut c_int, row_end: *mut c_int) -> c_int;
    pub fn swr_multi_set_rsqrt_table(m: *mut swr_multi, table: *const u32, mantissa_bits: c_int) -> c_int;
    pub fn swr_multi_upload_scene(m: *mut swr_multi, scene: *const swr_scene_desc) -> c_int;
    pub fn swr_multi_render(m: *mut swr_multi, camera: *const swr_camera, draws: *const swr_draw, ndraws: c_int) -> c_int;
    pub fn swr_multi_resolve(m: *mut swr_multi, exposure: f32, out_pixels: *mut u32) -> c_int;
    pub fn swr_multi_read_tile_luminance(m: *mut swr_multi, out_per_tile: *mut f32) -> c_int;
    pub fn swr_multi_get_stats(m: *mut swr_multi, out: *mut swr_frame_stats) -> c_int;
    pub fn swr_multi_synchronize(m: *mut swr_multi) -> c_int;
    pub fn swr_bake_brdf_lut(device: c_int, size: u32, out_texels: *mut u32) -> c_int;
    pub fn swr_bake_irradiance_sh4(device: c_int, cubemap_faces: *const u32, w: u32, h: u32, out12: *mut f32) -> c_int;
    pub fn swr_bake_prefilter_specular(device: c_int, cubemap_faces: *const u32, w: u32, h: u32, sample_count: u32, out_texels: *mut u32) -> c_int;
    pub fn swr_bake_sun_visibility(device: c_int, desc: *const swr_sunvis_desc, out_per_voxel: *mut f32) -> c_int;
    pub fn swr_bake_last_error() -> *const c_char;
    pub fn swr_device_pixels(ctx: *mut swr_ctx) -> *mut c_void;
    pub fn swr_device_keys(ctx: *mut swr_ctx) -> *mut c_void;
    pub fn swr_device_keys_bytes(ctx: *mut swr_ctx) -> usize;
    pub fn swr_cuda_stream(ctx: *mut swr_ctx) -> *mut c_void;
}

//! `extern "C"` declarations for libswr_b200.so (include/swr.h). Source only: this image has no Rust
//! toolchain, so this file is NOT compiled or tested here; the C++ mirror in
//! swraster-viewer_b200/host/swr_host.hpp exercises the same calls.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

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
pub struct swr_mesh_desc { pub first_primitive: u32, pub num_primitives: u32 }
#[repr(C)]
pub struct swr_node_desc { pub transform: [f32; 16], pub mesh_index: i32, pub bounding_sphere_world: [f32; 4] }
#[repr(C)]
pub struct swr_texture_desc {
    pub data: *const u32, pub ntexels: u32, pub width: u32, pub height: u32, pub texture_type: u32, pub max_mip_level: u32,
    pub mip_offsets: *const u32, pub mip_widths: *const u32, pub mip_heights: *const u32, pub array_stride: *const u32,
    pub wrap_s: u32, pub wrap_t: u32,
}
#[repr(C)]
pub struct swr_material_desc {
    pub base_color_factor: [f32; 4], pub metallic_factor: f32, pub roughness_factor: f32, pub emissive_factor: [f32; 3],
    pub occlusion_strength: f32, pub transmission: f32, pub alpha_cutoff: f32, pub flags: u32,
    pub base_color_texture: i32, pub metallic_roughness_texture: i32, pub normal_texture: i32,
    pub emissive_texture: i32, pub occlusion_texture: i32, pub transmission_texture: i32,
}
#[repr(C)]
pub struct swr_voxel_grid_desc { pub dims: [u32; 3], pub world_min: [f32; 3], pub world_max: [f32; 3], pub gi_sh4: *const f32 }
#[repr(C)]
pub struct swr_scene_desc {
    pub primitives: *const swr_primitive_desc, pub nprimitives: u32,
    pub meshes: *const swr_mesh_desc, pub nmeshes: u32,
    pub nodes: *const swr_node_desc, pub nnodes: u32,
    pub materials: *const swr_material_desc, pub nmaterials: u32,
    pub textures: *const swr_texture_desc, pub ntextures: u32,
    pub voxel_grid: swr_voxel_grid_desc,
    pub cubemap: i32, pub cubemap_specular: i32, pub brdf_lut: i32,
    pub light_direction: [f32; 3], pub light_color: [f32; 3],
}
#[repr(C)]
pub struct swr_camera {
    pub position: [f32; 4], pub view_matrix: [f32; 16], pub view_project_matrix: [f32; 16],
    pub skybox_matrix_transposed: [f32; 16], pub view_clip_planes: [[f32; 4]; 6],
    pub one_over_width: f32, pub one_over_height: f32, pub reserved: [f32; 2],
}
#[repr(C)]
pub struct swr_draw { pub model: [f32; 16], pub mvp: [f32; 16], pub primitive: u32, pub flags: u32, pub first_triangle: u32, pub reserved: u32 }
#[repr(C)]
pub struct swr_ctx { _private: [u8; 0] }

#[link(name = "swr_b200")]
extern "C" {
    pub fn swr_create(width: c_int, height: c_int, device: c_int) -> *mut swr_ctx;
    pub fn swr_destroy(ctx: *mut swr_ctx);
    pub fn swr_last_error(ctx: *const swr_ctx) -> *const c_char;
    pub fn swr_set_rsqrt_table(ctx: *mut swr_ctx, table: *const u32, mantissa_bits: c_int) -> c_int;
    pub fn swr_set_tile_rows(ctx: *mut swr_ctx, row_begin: c_int, row_end: c_int) -> c_int;
    pub fn swr_upload_scene(ctx: *mut swr_ctx, scene: *const swr_scene_desc) -> c_int;
    pub fn swr_render(ctx: *mut swr_ctx, camera: *const swr_camera, draws: *const swr_draw, ndraws: c_int, shade: c_int) -> c_int;
    pub fn swr_resolve(ctx: *mut swr_ctx, exposure: f32, out_pixels: *mut u32) -> c_int;
    pub fn swr_read_tile_luminance(ctx: *mut swr_ctx, out_per_tile: *mut f32) -> c_int;
    pub fn swr_device_pixels(ctx: *mut swr_ctx) -> *mut c_void;
    // sort-first frame assembly over NVLink peer memory (one process per GPU): see include/swr.h
    pub fn swr_peer_export(ctx: *mut swr_ctx, handle_out: *mut u8) -> c_int; // SWR_PEER_HANDLE_BYTES = 64
    pub fn swr_peer_open(ctx: *mut swr_ctx, handle: *const u8) -> c_int;
    pub fn swr_resolve_peer(ctx: *mut swr_ctx, exposure: f32, frame: u32) -> c_int;
    pub fn swr_peer_collect(ctx: *mut swr_ctx, frame: u32, contributors: c_int) -> c_int;
    pub fn swr_peer_release(ctx: *mut swr_ctx, frame: u32) -> c_int;
}

//! Drop-in body for `src/renderer.rs` of swraster-viewer: same public API
//! (`Renderer::new`, `render_scene`, `update_auto_exposure`, `blit_to_buffer`, `RenderBuffer`),
//! device work behind the C ABI. Source only (no Rust toolchain in this image — not compiled here).
use crate::ffi::*;
use crate::rendercamera::RenderCamera;
use crate::scene::{BoundingSphere, Scene};
use glam::{Mat4, Vec4};
use ordered_float::OrderedFloat;

pub struct Renderer {
    ctx: *mut swr_ctx,
    uploaded: *const Scene, // Scene is immutable after load: upload on first sight
    draws: Vec<swr_draw>,
    tile_luminance: Vec<f32>,
    auto_exposure: f32, auto_exposure_target: f32, auto_exposure_ev: f32,
}
unsafe impl Send for Renderer {} // one call at a time per context, any thread (main.rs:536 holds a Mutex)

impl Renderer {
    pub fn new(width: i32, height: i32) -> Self {
        let ctx = unsafe { swr_create(width, height, 0) };
        assert!(!ctx.is_null(), "swr_create failed (no CPU fallback)");
        // math.rs:34-39: hand the host's _mm_rsqrt_ps table to the device so normalize() matches this CPU
        let (table, bits) = probe_host_rsqrt_table();
        if bits > 0 { unsafe { swr_set_rsqrt_table(ctx, table.as_ptr(), bits) }; }
        let tiles = (((width + 63) / 64) * ((height + 63) / 64)) as usize;
        Self { ctx, uploaded: std::ptr::null(), draws: vec![], tile_luminance: vec![1.0; tiles],
               auto_exposure: 2.0, auto_exposure_target: 2.0, auto_exposure_ev: 1.0 }
    }

    pub fn render_scene(&mut self, scene: &Scene, camera: &RenderCamera) {
        if self.uploaded != scene as *const Scene {
            let desc = scene_desc(scene); // flat POD view of the Vec<..> fields of scene.rs:65-121
            check(self.ctx, unsafe { swr_upload_scene(self.ctx, &desc.pod) });
            self.uploaded = scene;
        }
        // renderer.rs:357-367: stable sort of node indices by squared distance to the camera
        let mut order: Vec<usize> = (0..scene.nodes.len()).collect();
        order.sort_by_key(|&i| OrderedFloat((camera.position - scene.nodes[i].bounding_sphere_world.center).length_squared()));
        // renderer.rs:369-468: mvp, sphere/frustum classification, opaque primitives in mesh order
        self.draws.clear();
        let mut first_triangle = 0u32;
        for i in order {
            let node = &scene.nodes[i];
            let Some(mesh_index) = node.mesh_index else { continue };
            let mvp = camera.view_project_matrix * node.transform;
            let mesh = &scene.meshes[mesh_index];
            for &p in &mesh.primitives_opaque {
                let prim = &mesh.primitives[p];
                let sphere = node.transform * &prim.bounding_sphere;
                let flags = match test_sphere_frustum(&sphere, camera) { Frustum::Outside => continue, Frustum::Inside => 0, Frustum::Intersecting => 1 };
                self.draws.push(swr_draw { model: node.transform.to_cols_array(), mvp: mvp.to_cols_array(),
                                           primitive: desc_primitive_index(scene, mesh_index, p), flags, first_triangle, reserved: 0 });
                first_triangle += (prim.indices.len() / 3) as u32;
            }
        }
        let cam = camera_pod(camera);
        check(self.ctx, unsafe { swr_render(self.ctx, &cam, self.draws.as_ptr(), self.draws.len() as i32, 1) });
    }

    pub fn update_auto_exposure(&mut self, delta_time: f32) {
        check(self.ctx, unsafe { swr_read_tile_luminance(self.ctx, self.tile_luminance.as_mut_ptr()) });
        /* renderer.rs:264-289 unchanged, reading self.tile_luminance instead of tile.center_luminance */
    }

    pub fn blit_to_buffer(&self, buffer: &mut RenderBuffer) {
        check(self.ctx, unsafe { swr_resolve(self.ctx, self.auto_exposure, buffer.pixels.as_mut_ptr()) });
    }
}
impl Drop for Renderer { fn drop(&mut self) { unsafe { swr_destroy(self.ctx) } } }

fn check(ctx: *mut swr_ctx, rc: i32) {
    if rc != 0 { panic!("swr: {}", unsafe { std::ffi::CStr::from_ptr(swr_last_error(ctx)) }.to_string_lossy()); } // the reference panics on this path
}

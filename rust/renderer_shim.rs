//! Drop-in replacement for `src/renderer.rs` of swraster-viewer: the same public API
//! (`Renderer::new` :165, `render_scene` :201, `update_auto_exposure` :258, `blit_to_buffer` :293, `RenderBuffer` :24-28,
//! :78-96), with the per-frame device work behind the C ABI of include/swr.h (`crate::ffi`, rust/ffi.rs).
//! `Scene`, `RenderCamera`, `main.rs` stay untouched; `tilerasterizer.rs`, `shader.rs`, `bumpqueue.rs` drop out of the build.
//!
//! What stays on the host, unchanged from the reference: the node sort (renderer.rs:357-367), the per-primitive
//! sphere / frustum classification and `mvp = view_project * model` (renderer.rs:369-468, scene.rs:53-63), and the
//! auto-exposure metering maths (renderer.rs:258-290). Opaque primitives are submitted before translucent ones per mesh
//! (renderer.rs:386-420); translucent draws carry SWR_DRAW_TRANSLUCENT and their own running triangle count.
//!
//! Devices: `SWR_DEVICES=0,1,2,3` (environment) puts ONE Renderer over several GPUs of the box (sort-first,
//! swr_multi_*); unset = device 0.
//!
//! Source only: this image has no Rust toolchain, so the file is not compiled here. tests/test_rust_shim.py checks that
//! every `swr_*` call below exists in rust/ffi.rs with the argument count used here, and that ffi.rs matches include/swr.h.
use crate::ffi::*;
use crate::rendercamera::RenderCamera;
use crate::scene::{BoundingSphere, Material, Scene};
use crate::texture::{Texture, TextureAndSampler, TextureType, WrapMode};
use glam::Mat4;
use ordered_float::OrderedFloat;
use std::os::raw::c_int;
use std::sync::Arc;

const TILE_SIZE: i32 = 64;
const DEFAULT_EXPOSURE: f32 = 2.0;
const AUTO_EXPOSURE_MID_GRAY_POST_TONEMAP: f32 = 0.45;
const AUTO_EXPOSURE_MIN: f32 = 0.05;
const AUTO_EXPOSURE_MAX: f32 = 32.0;
const AUTO_EXPOSURE_TRIM_FRACTION: f32 = 0.10;
const AUTO_EXPOSURE_TIME_CONSTANT_SECONDS: f32 = 1.0;
const TONEMAP_K: f32 = 0.2; // util.rs:37-47

pub struct RenderBuffer<'a> {
    pub width: usize,
    pub height: usize,
    pub pixels: &'a mut [u32],
}

impl<'a> RenderBuffer<'a> {
    pub fn new(width: usize, height: usize, pixels: &'a mut [u32]) -> Self {
        Self { width, height, pixels }
    }
    pub fn clear(&mut self) {
        self.pixels.fill(0);
    }
    pub fn set_pixel(&mut self, x: usize, y: usize, color: u32) {
        if x < self.width && y < self.height {
            self.pixels[y * self.width + x] = color;
        }
    }
}

#[derive(PartialEq)]
enum FrustumTestResult {
    Inside,
    Outside,
    Intersecting,
}

// renderer.rs:130-142, unchanged
fn test_sphere_frustum(sphere: &BoundingSphere, camera: &RenderCamera) -> FrustumTestResult {
    let center_view = camera.view_matrix * sphere.center.extend(1.0);
    let mut result = FrustumTestResult::Inside;
    for plane in &camera.view_clip_planes {
        let distance = plane.dot(center_view);
        if distance < -sphere.radius {
            return FrustumTestResult::Outside;
        } else if distance < sphere.radius {
            result = FrustumTestResult::Intersecting;
        }
    }
    result
}

// util.rs:43-47
fn tonemap_inverse_scalar(mapped: f32) -> f32 {
    let y = mapped.clamp(0.0, 1.0);
    let denom = (1.0 + TONEMAP_K - y).max(1e-6);
    (y * TONEMAP_K) / denom
}

/// math.rs:34-39: on x86-64 `rsqrt_vec` IS `_mm_rsqrt_ps`, whose value depends only on the exponent parity and the top K
/// mantissa bits of the input — a table that belongs to the CPU this program runs on. Probe it (K = 8..16, exhaustive
/// check of the accepted K) and hand it to the device so the CUDA shader normalises exactly like this host would.
/// Returns (table, K); K = 0 when the estimate has no such structure here (the device then uses rsqrtf()).
fn probe_host_rsqrt_table() -> (Vec<u32>, c_int) {
    #[cfg(target_arch = "x86_64")]
    {
        use std::arch::x86_64::{_mm_cvtss_f32, _mm_rsqrt_ss, _mm_set_ss};
        let rs = |b: u32| -> u32 { unsafe { _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(f32::from_bits(b)))).to_bits() } };
        for k in 8..=16u32 {
            let n = 1usize << k;
            let mut table = vec![0u32; 2 * n];
            for p in 0..2u32 {
                for i in 0..n as u32 {
                    table[(p as usize) * n + i as usize] = rs(((127 + p) << 23) | (i << (23 - k)));
                }
            }
            let bucket_ok = |step: usize| -> bool {
                (0..2u32).all(|p| (0..(1u32 << 23)).step_by(step).all(|m| rs(((127 + p) << 23) | m) == table[(p as usize) * n + (m >> (23 - k)) as usize]))
            };
            if !bucket_ok(61) || !bucket_ok(1) {
                continue;
            }
            // exponent scaling: +2 in the exponent halves the result
            let scale_ok = (1..253u32).step_by(2).all(|e| (0..(1u32 << 23)).step_by(4099).all(|m| rs((e << 23) | m).wrapping_sub(rs(((e + 2) << 23) | m)) == 1 << 23));
            if scale_ok {
                return (table, k as c_int);
            }
        }
    }
    (Vec::new(), 0)
}

fn wrap_code(w: &WrapMode) -> u32 {
    match w {
        WrapMode::Repeat => SWR_WRAP_REPEAT,
        WrapMode::MirroredRepeat => SWR_WRAP_MIRRORED_REPEAT,
        WrapMode::ClampToEdge => SWR_WRAP_CLAMP_TO_EDGE,
    }
}

fn type_code(t: &TextureType) -> u32 {
    match t {
        TextureType::SRGB => SWR_TEX_SRGB,
        TextureType::Normal => SWR_TEX_NORMAL,
        TextureType::MetallicRoughness => SWR_TEX_METALLIC_ROUGHNESS,
        TextureType::Cubemap => SWR_TEX_CUBEMAP,
        TextureType::Linear => SWR_TEX_LINEAR,
    }
}

/// Flat POD view of a `Scene` (scene.rs:65-121): the tables are owned here, every pointer inside them points into the
/// `Vec`s of the borrowed Scene, which is immutable after load and outlives the renderer (main.rs:283-290, :312).
struct ScenePod {
    primitives: Vec<swr_primitive_desc>,
    meshes: Vec<swr_mesh_desc>,
    nodes: Vec<swr_node_desc>,
    materials: Vec<swr_material_desc>,
    textures: Vec<swr_texture_desc>,
    texture_keys: Vec<(*const Texture, u32, u32)>, // (Arc pointer, wrap_s, wrap_t) of textures[i]
    mesh_first_primitive: Vec<u32>,
}

impl ScenePod {
    fn texture_index(&mut self, t: &TextureAndSampler) -> i32 {
        let key = (Arc::as_ptr(&t.texture), wrap_code(&t.sampler.wrap_s), wrap_code(&t.sampler.wrap_t));
        if let Some(i) = self.texture_keys.iter().position(|k| *k == key) {
            return i as i32;
        }
        let tex: &Texture = &t.texture;
        self.textures.push(swr_texture_desc {
            data: tex.data.as_ptr(),
            ntexels: tex.data.len() as u32,
            width: tex.width,
            height: tex.height,
            texture_type: type_code(&tex.texture_type),
            max_mip_level: tex.max_mip_level,
            mip_offsets: tex.mip_offsets.as_ptr(),
            mip_widths: tex.mip_widths.as_ptr(),
            mip_heights: tex.mip_heights.as_ptr(),
            array_stride: tex.array_stride.as_ptr(),
            wrap_s: key.1,
            wrap_t: key.2,
        });
        self.texture_keys.push(key);
        (self.textures.len() - 1) as i32
    }

    fn optional_texture(&mut self, t: &Option<TextureAndSampler>) -> i32 {
        match t {
            Some(t) => self.texture_index(t),
            None => -1,
        }
    }

    fn material(&mut self, m: &Material) -> swr_material_desc {
        swr_material_desc {
            base_color_factor: m.base_color_factor.to_array(),
            metallic_factor: m.metallic_factor,
            roughness_factor: m.roughness_factor,
            emissive_factor: m.emissive_factor.to_array(),
            occlusion_strength: m.occlusion_strength,
            transmission: m.transmission,
            alpha_cutoff: m.alpha_cutoff_vec.x, // a splat (scene.rs:706-707)
            flags: (if m.is_alpha_tested { SWR_MAT_ALPHA_TESTED } else { 0 }) | (if m.is_translucent { SWR_MAT_TRANSLUCENT } else { 0 }),
            base_color_texture: self.optional_texture(&m.base_color_texture),
            metallic_roughness_texture: self.optional_texture(&m.metallic_roughness_texture),
            normal_texture: self.optional_texture(&m.normal_texture),
            emissive_texture: self.optional_texture(&m.emissive_texture),
            occlusion_texture: self.optional_texture(&m.occlusion_texture),
            transmission_texture: self.optional_texture(&m.transmission_texture),
        }
    }

    /// Builds the tables and returns them with the scene-level descriptor that points at them.
    fn new(scene: &Scene) -> (Self, swr_scene_desc) {
        let mut pod = ScenePod { primitives: vec![], meshes: vec![], nodes: vec![], materials: vec![], textures: vec![], texture_keys: vec![], mesh_first_primitive: vec![] };
        for mesh in &scene.meshes {
            pod.mesh_first_primitive.push(pod.primitives.len() as u32);
            pod.meshes.push(swr_mesh_desc { first_primitive: pod.primitives.len() as u32, num_primitives: mesh.primitives.len() as u32 });
            for p in &mesh.primitives {
                let c = p.bounding_sphere.center;
                pod.primitives.push(swr_primitive_desc {
                    positions: p.positions.as_ptr() as *const f32,
                    normals: p.normals.as_ptr() as *const f32,
                    tangents: p.tangents.as_ptr() as *const f32,
                    texcoords: p.texcoords.as_ptr() as *const f32,
                    indices: p.indices.as_ptr(),
                    nverts: p.positions.len() as u32,
                    nindices: p.indices.len() as u32,
                    material_index: p.material_index as u32,
                    bounding_sphere: [c.x, c.y, c.z, p.bounding_sphere.radius],
                });
            }
        }
        for n in &scene.nodes {
            let c = n.bounding_sphere_world.center;
            pod.nodes.push(swr_node_desc {
                transform: n.transform.to_cols_array(),
                mesh_index: n.mesh_index.map(|i| i as i32).unwrap_or(-1),
                bounding_sphere_world: [c.x, c.y, c.z, n.bounding_sphere_world.radius],
            });
        }
        for m in &scene.materials {
            let d = pod.material(m);
            pod.materials.push(d);
        }
        let cubemap = pod.texture_index(&scene.cubemap);
        let cubemap_specular = pod.texture_index(&scene.cubemap_specular);
        let brdf_lut = pod.texture_index(&scene.brdf_lut);
        let (gw, gh, gd) = scene.voxel_grid.dimensions();
        let desc = swr_scene_desc {
            primitives: pod.primitives.as_ptr(),
            nprimitives: pod.primitives.len() as u32,
            meshes: pod.meshes.as_ptr(),
            nmeshes: pod.meshes.len() as u32,
            nodes: pod.nodes.as_ptr(),
            nnodes: pod.nodes.len() as u32,
            materials: pod.materials.as_ptr(),
            nmaterials: pod.materials.len() as u32,
            textures: pod.textures.as_ptr(),
            ntextures: pod.textures.len() as u32,
            voxel_grid: swr_voxel_grid_desc {
                dims: [gw as u32, gh as u32, gd as u32],
                world_min: scene.voxel_grid.world_min().to_array(),
                world_max: scene.voxel_grid.world_max().to_array(),
                gi_sh4: scene.voxel_grid.gi_sh4().as_ptr() as *const f32, // Vec<[Vec4; 4]>: 16 floats per voxel
            },
            cubemap,
            cubemap_specular,
            brdf_lut,
            light_direction: scene.light.direction.to_array(),
            light_color: scene.light.color.to_array(),
        };
        (pod, desc)
    }
}

fn camera_pod(camera: &RenderCamera) -> swr_camera {
    let p = camera.position;
    let mut planes = [[0.0f32; 4]; 6];
    for (dst, src) in planes.iter_mut().zip(camera.view_clip_planes.iter()) {
        *dst = src.to_array();
    }
    swr_camera {
        position: [p.x, p.y, p.z, 0.0],
        view_matrix: camera.view_matrix.to_cols_array(),
        view_project_matrix: camera.view_project_matrix.to_cols_array(),
        skybox_matrix_transposed: camera.skybox_matrix_transposed.to_cols_array(),
        view_clip_planes: planes,
        one_over_width: camera.one_over_width,
        one_over_height: camera.one_over_height,
        reserved: [0.0; 2],
    }
}

enum Backend {
    Single(*mut swr_ctx),
    Multi(*mut swr_multi),
}

pub struct Renderer {
    backend: Backend,
    width: i32,
    height: i32,
    uploaded: *const Scene, // Scene is immutable after load: upload on first sight, keyed by address
    pod: Option<ScenePod>,
    draws: Vec<swr_draw>,
    nodes_by_distance: Vec<usize>,
    tile_luminance: Vec<f32>,
    auto_exposure: f32,
    auto_exposure_target: f32,
    auto_exposure_ev: f32,
    // Fixed-exposure shading (include/swr.h swr_set_fixed_exposure): the exposure the frame rendered last was shaded with
    // (0.0 = HDR), whether update_auto_exposure moved the exposure the last time it ran, and that frame's camera block
    // (shading the same visibility buffer again needs it). Cell: blit_to_buffer takes &self, like the reference's.
    frame_fixed: std::cell::Cell<f32>,
    meter_live: bool,
    last_cam: Option<swr_camera>,
}
// App holds Mutex<Renderer> and calls from a rayon worker (main.rs:526-543): one call at a time, any thread.
unsafe impl Send for Renderer {}

impl Renderer {
    pub fn new(width: i32, height: i32) -> Self {
        let devices: Vec<c_int> = std::env::var("SWR_DEVICES").ok().map(|s| s.split(',').filter_map(|d| d.trim().parse().ok()).collect()).unwrap_or_default();
        let (table, bits) = probe_host_rsqrt_table();
        let table_ptr = if bits > 0 { table.as_ptr() } else { std::ptr::null() };
        let backend = if devices.len() > 1 {
            let m = unsafe { swr_multi_create(width, height, devices.as_ptr(), devices.len() as c_int, SWR_MULTI_SORT_FIRST) };
            if m.is_null() {
                panic!("swr_multi_create: {}", unsafe { std::ffi::CStr::from_ptr(swr_multi_last_error(std::ptr::null())) }.to_string_lossy());
            }
            let rc = unsafe { swr_multi_set_rsqrt_table(m, table_ptr, bits) };
            assert!(rc == SWR_OK, "swr_multi_set_rsqrt_table failed");
            Backend::Multi(m)
        } else {
            let device = devices.first().copied().unwrap_or(0);
            let ctx = unsafe { swr_create(width, height, device) };
            if ctx.is_null() {
                // there is no CPU fallback: the reference's behaviour on an unusable renderer is a panic
                panic!("swr_create: {}", unsafe { std::ffi::CStr::from_ptr(swr_last_error(std::ptr::null())) }.to_string_lossy());
            }
            let rc = unsafe { swr_set_rsqrt_table(ctx, table_ptr, bits) };
            assert!(rc == SWR_OK, "swr_set_rsqrt_table failed");
            Backend::Single(ctx)
        };
        let tiles_x = (width + TILE_SIZE - 1) / TILE_SIZE;
        let tiles_y = (height + TILE_SIZE - 1) / TILE_SIZE;
        Self {
            backend,
            width,
            height,
            uploaded: std::ptr::null(),
            pod: None,
            draws: Vec::new(),
            nodes_by_distance: Vec::new(),
            tile_luminance: vec![1.0; (tiles_x * tiles_y) as usize], // tilerasterizer: center_luminance starts at 1.0
            auto_exposure: DEFAULT_EXPOSURE,
            auto_exposure_target: DEFAULT_EXPOSURE,
            auto_exposure_ev: DEFAULT_EXPOSURE.log2(),
            frame_fixed: std::cell::Cell::new(0.0),
            meter_live: false,
            last_cam: None,
        }
    }

    fn check(&self, rc: c_int, what: &str) {
        if rc != SWR_OK {
            let msg = match self.backend {
                Backend::Single(ctx) => unsafe { std::ffi::CStr::from_ptr(swr_last_error(ctx)) },
                Backend::Multi(m) => unsafe { std::ffi::CStr::from_ptr(swr_multi_last_error(m)) },
            };
            panic!("{}: {}", what, msg.to_string_lossy()); // the reference panics on this path (slice indexing)
        }
    }

    // Main render function for the scene (renderer.rs:201)
    pub fn render_scene(&mut self, scene: &Scene, camera: &RenderCamera) {
        if self.uploaded != scene as *const Scene {
            let (pod, desc) = ScenePod::new(scene);
            let rc = match self.backend {
                Backend::Single(ctx) => unsafe { swr_upload_scene(ctx, &desc) },
                Backend::Multi(m) => unsafe { swr_multi_upload_scene(m, &desc) },
            };
            self.check(rc, "swr_upload_scene");
            self.pod = Some(pod); // the device has its own copy now; the tables are kept for the primitive index map
            self.uploaded = scene;
        }
        // renderer.rs:357-367: stable sort of the node indices by squared distance camera -> node sphere centre
        self.nodes_by_distance.clear();
        self.nodes_by_distance.extend(0..scene.nodes.len());
        self.nodes_by_distance.sort_by_key(|&i| OrderedFloat((camera.position - scene.nodes[i].bounding_sphere_world.center).length_squared()));
        // renderer.rs:369-468: per node mvp; per mesh the opaque primitives, then the translucent ones; per primitive the
        // sphere / frustum classification picks the clipping or the non-clipping instantiation, Outside submits nothing
        let pod = self.pod.as_ref().unwrap();
        self.draws.clear();
        let (mut first_triangle, mut first_triangle_translucent) = (0u32, 0u32);
        for &ni in &self.nodes_by_distance {
            let node = &scene.nodes[ni];
            let Some(mesh_index) = node.mesh_index else { continue };
            let model: Mat4 = node.transform;
            let mvp: Mat4 = camera.view_project_matrix * model;
            let mesh = &scene.meshes[mesh_index];
            for (list, translucent) in [(&mesh.primitives_opaque, false), (&mesh.primitives_translucent, true)] {
                for &p in list.iter() {
                    let prim = &mesh.primitives[p];
                    let sphere_world = model * &prim.bounding_sphere; // scene.rs:53-63: radius * mean axis length
                    let clip = match test_sphere_frustum(&sphere_world, camera) {
                        FrustumTestResult::Outside => continue,
                        FrustumTestResult::Inside => 0,
                        FrustumTestResult::Intersecting => SWR_DRAW_CLIP,
                    };
                    let ntris = (prim.indices.len() / 3) as u32;
                    let counter = if translucent { &mut first_triangle_translucent } else { &mut first_triangle };
                    self.draws.push(swr_draw {
                        model: model.to_cols_array(),
                        mvp: mvp.to_cols_array(),
                        primitive: pod.mesh_first_primitive[mesh_index] + p as u32,
                        flags: clip | if translucent { SWR_DRAW_TRANSLUCENT } else { 0 },
                        first_triangle: *counter,
                        reserved: 0,
                    });
                    *counter += ntris;
                }
            }
        }
        let cam = camera_pod(camera);
        // renderer.rs:258-290 is the only writer of the exposure: while it leaves it alone, blit_to_buffer will use the value
        // held now, and shading can pack RGBA8 with it (no HDR round trip); frame_exposure() below covers the other case
        let fixed = if self.meter_live { 0.0 } else { self.auto_exposure };
        if let Backend::Single(ctx) = self.backend {
            let rc = unsafe { swr_set_fixed_exposure(ctx, fixed) };
            self.check(rc, "swr_set_fixed_exposure");
            self.frame_fixed.set(fixed);
        }
        self.last_cam = Some(cam);
        let rc = match self.backend {
            Backend::Single(ctx) => unsafe { swr_render(ctx, &cam, self.draws.as_ptr(), self.draws.len() as c_int, 1) },
            Backend::Multi(m) => unsafe { swr_multi_render(m, &cam, self.draws.as_ptr(), self.draws.len() as c_int) },
        };
        self.check(rc, "swr_render");
    }

    // renderer.rs:258-290 with the tiles' center_luminance read back from the device
    pub fn update_auto_exposure(&mut self, delta_time: f32) {
        let sample_count = self.tile_luminance.len();
        if sample_count == 0 {
            return;
        }
        let rc = match self.backend {
            Backend::Single(ctx) => unsafe { swr_read_tile_luminance(ctx, self.tile_luminance.as_mut_ptr()) },
            Backend::Multi(m) => unsafe { swr_multi_read_tile_luminance(m, self.tile_luminance.as_mut_ptr()) },
        };
        self.check(rc, "swr_read_tile_luminance");

        let mut tile_log_luminance = vec![0.0f32; sample_count];
        for (sample, lum) in tile_log_luminance.iter_mut().zip(self.tile_luminance.iter()) {
            *sample = lum.max(1e-4).log2();
        }
        let samples = &mut tile_log_luminance[..];
        samples.sort_unstable_by(|a, b| a.total_cmp(b));
        let trim_count = ((sample_count as f32) * AUTO_EXPOSURE_TRIM_FRACTION).floor() as usize;
        let trim_count = trim_count.min((sample_count - 1) / 2);
        let trimmed = &samples[trim_count..(sample_count - trim_count)];
        let mean_log_luminance = trimmed.iter().copied().sum::<f32>() / trimmed.len() as f32;

        let meter_key = tonemap_inverse_scalar(AUTO_EXPOSURE_MID_GRAY_POST_TONEMAP).max(1e-4);
        let target_ev = meter_key.log2() - mean_log_luminance;
        let target = (2.0f32.powf(target_ev)).clamp(AUTO_EXPOSURE_MIN, AUTO_EXPOSURE_MAX);
        self.auto_exposure_target = target;

        let target_ev = self.auto_exposure_target.log2();
        let tau = AUTO_EXPOSURE_TIME_CONSTANT_SECONDS.max(1e-4);
        let alpha = 1.0 - (-(delta_time.max(0.0) / tau)).exp();
        self.auto_exposure_ev += (target_ev - self.auto_exposure_ev) * alpha;
        let moved = 2.0f32.powf(self.auto_exposure_ev);
        self.meter_live = moved != self.auto_exposure; // moving: the next frames are shaded to HDR until it holds still
        self.auto_exposure = moved;
    }

    // The frame rendered last must be resolvable with `exposure`: one that was packed with another exposure is shaded again
    // from its visibility buffer, to HDR (swr::Renderer::frame_exposure in host/swr_host.hpp is the tested twin of this).
    fn frame_exposure(&self, exposure: f32) {
        let fixed = self.frame_fixed.get();
        if fixed == 0.0 || fixed == exposure {
            return;
        }
        if let (Backend::Single(ctx), Some(cam)) = (&self.backend, self.last_cam.as_ref()) {
            let rc = unsafe { swr_set_fixed_exposure(*ctx, 0.0) };
            self.check(rc, "swr_set_fixed_exposure");
            let rc = unsafe { swr_shade(*ctx, cam) };
            self.check(rc, "swr_shade");
            self.frame_fixed.set(0.0);
        }
    }

    // Copy the frame to the backbuffer (renderer.rs:293): exposure, tonemap, RGBA8 pack on the device, then W*H u32 to the host
    pub fn blit_to_buffer(&self, buffer: &mut RenderBuffer) {
        assert!(buffer.width == self.width as usize && buffer.pixels.len() >= (self.width * self.height) as usize, "RenderBuffer size mismatch");
        self.frame_exposure(self.auto_exposure);
        let rc = match self.backend {
            Backend::Single(ctx) => unsafe { swr_resolve(ctx, self.auto_exposure, buffer.pixels.as_mut_ptr()) },
            Backend::Multi(m) => unsafe { swr_multi_resolve(m, self.auto_exposure, buffer.pixels.as_mut_ptr()) },
        };
        self.check(rc, "swr_resolve");
    }
}

impl Drop for Renderer {
    fn drop(&mut self) {
        match self.backend {
            Backend::Single(ctx) => unsafe { swr_destroy(ctx) },
            Backend::Multi(m) => unsafe { swr_multi_destroy(m) },
        }
    }
}

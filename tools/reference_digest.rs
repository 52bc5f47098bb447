//! tools/reference_digest.rs — pins the CPU oracle (oracle/oracle.cpp) against the REAL swraster-viewer.
//!
//! The build image of the CUDA port has no Rust toolchain, so the oracle's parity with the reference is unpinned there
//!. Anyone with `cargo` closes that gap in a few minutes:
//!
//!   1. `python tools/export_reference_inputs.py out/`      (CUDA repo) writes out/c1.gltf, c2.gltf, c3.gltf and prints
//!      the environment lines for step 3;
//!   2. paste PART A into `impl Renderer` in src/renderer.rs and PART B at the end of `load_scene` in src/main.rs
//!      (just before the scene is handed to the app), in a checkout of mdesmedt/swraster-viewer;
//!   3. `RAYON_NUM_THREADS=1 SWR_DIGEST_OUT=c3.json SWR_DIGEST_CAMERA="..." cargo run --release -- out/c3.gltf`
//!      (RAYON_NUM_THREADS=1 makes the per-tile queue order the serial submission order the oracle models);
//!   4. copy the JSON files to tests/golden/reference_digests/ of the CUDA repo: tests/test_reference_digest.py then
//!      compares them with the oracle tile by tile (visibility digest + covered-lane count; the colour digest is
//!      reported, not required: it depends on the host CPU's _mm_rsqrt_ps).
//!
//! Digest per tile (the same bytes as orc_tile_digests in oracle/oracle.cpp): FNV-1a-64 over, per quad and lane,
//! depth bits, and for lanes with depth != +INF also packet_index, bary1 bits, bary2 bits.

// ---------------------------------------------------------------- PART A: src/renderer.rs, inside `impl Renderer`
/*
    pub fn visibility_digests(&self) -> Vec<[u64; 3]> {
        fn fnv(h: &mut u64, w: u32) {
            for k in 0..4 {
                *h ^= ((w >> (8 * k)) & 0xFF) as u64;
                *h = h.wrapping_mul(0x100000001b3);
            }
        }
        self.tiles
            .iter()
            .map(|t| {
                let (mut hv, mut hc, mut covered) = (0xcbf29ce484222325u64, 0xcbf29ce484222325u64, 0u64);
                for q in 0..t.depth.len() {
                    let (d, p, b1, b2) = (t.depth[q].to_array(), t.packet_index[q].to_array(), t.bary1[q].to_array(), t.bary2[q].to_array());
                    for l in 0..4 {
                        fnv(&mut hv, d[l].to_bits());
                        if d[l].to_bits() != 0x7F80_0000 {
                            covered += 1;
                            fnv(&mut hv, p[l]);
                            fnv(&mut hv, b1[l].to_bits());
                            fnv(&mut hv, b2[l].to_bits());
                        }
                    }
                    for c in [t.color[q].x, t.color[q].y, t.color[q].z] {
                        for v in c.to_array() {
                            fnv(&mut hc, v.to_bits());
                        }
                    }
                }
                [hv, covered, hc]
            })
            .collect()
    }
*/

// ---------------------------------------------------------------- PART B: src/main.rs, end of `load_scene`
// (`scene` is the loaded Scene with voxel grid and cubemaps attached; the camera block replaces the default camera)
/*
    if let Ok(out_path) = std::env::var("SWR_DIGEST_OUT") {
        // SWR_DIGEST_CAMERA = "px py pz  lx ly lz  mouse_dy  fov  far  width height": RenderCamera::new towards the LEVEL
        // target (lx, ly, lz), then rotate_mouse(0, mouse_dy) for the pitch — exactly how the CUDA repo builds its cameras
        let v: Vec<f32> = std::env::var("SWR_DIGEST_CAMERA").expect("SWR_DIGEST_CAMERA").split_whitespace().map(|x| x.parse().unwrap()).collect();
        let (w, h) = (v[9] as i32, v[10] as i32);
        let mut cam = RenderCamera::new(Vec3A::new(v[0], v[1], v[2]), Vec3A::new(v[3], v[4], v[5]), v[7], w as f32, h as f32, v[8]);
        cam.rotate_mouse(glam::Vec2::new(0.0, v[6]));
        cam.update_matrices();
        let mut renderer = Renderer::new(w, h); // fresh: barycentric / id buffers start at zero
        renderer.render_scene(&scene, &cam);
        let rows: Vec<String> = renderer.visibility_digests().iter().map(|d| format!("[\"{:016x}\", {}, \"{:016x}\"]", d[0], d[1], d[2])).collect();
        std::fs::write(&out_path, format!("{{\"width\": {}, \"height\": {}, \"tiles\": [{}]}}\n", w, h, rows.join(", "))).unwrap();
        println!("wrote {} tile digests to {}", rows.len(), out_path);
        std::process::exit(0);
    }
*/

"""Quick per-phase timing of the CUDA path on the BASELINE configs (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes

which = sys.argv[1:] or ["c1", "c2", "c3"]
for name in which:
    t0 = time.time()
    if name == "c1":
        sc, spec = scenes.scene_c1_sphere(voxel_dim=128, cube_size=256); W, H = 1920, 1080
    elif name == "c2":
        sc, spec = scenes.scene_c2_terrain(voxel_dim=128, cube_size=256); W, H = 1920, 1080
    elif name == "c3":
        sc, spec = scenes.scene_c3_instanced(voxel_dim=128, cube_size=256); W, H = 3840, 2160
    elif name == "c4":
        sc, spec = scenes.scene_c4_micro(voxel_dim=128, cube_size=256); W, H = 3840, 2160
    elif name == "c4s":
        sc, spec = scenes.scene_c4_micro(2001, voxel_dim=128, cube_size=256); W, H = 3840, 2160
    print(f"[{name}] scene built in {time.time()-t0:.1f}s: {sc.total_triangles} tris", flush=True)
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    buf = swr.RenderBuffer(W, H, pinned=True)
    for i in range(6):
        t0 = time.time()
        r.render_scene(sc, cam)
        r.blit_to_buffer(buf)
        dt = (time.time() - t0) * 1e3
        st = r.stats()
        if i in (0, 5):
            print(f"[{name}] frame {i}: wall {dt:.2f} ms | setup+bin {st['ms_setup_bin']:.3f} raster {st['ms_raster']:.3f} shade {st['ms_shade']:.3f} "
                  f"resolve {st['ms_resolve']:.3f} | T={st['triangles_submitted']} binned={st['triangles_binned']} clipped={st['triangles_clipped']} R={st['tile_refs']}", flush=True)
    r.close()

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
sc, spec = scenes.scene_c3_instanced(voxel_dim=16, cube_size=32)
W, H = 3840, 2160
cam = swr.RenderCamera.from_spec(spec, W, H)
for rows in [(0, 34), (0, 7), (0, 2)]:
    r = swr.Renderer(W, H); r.set_tile_rows(*rows)
    for i in range(6):
        r.render_scene(sc, cam); r.resolve_device_only(2.0); r.synchronize()
    st = r.stats(); refs, cyc = r.read_tile_costs()
    print(rows, f"draws {r.num_draws} setup {st['ms_setup_bin']:.3f} raster {st['ms_raster']:.3f} shade {st['ms_shade']:.3f} refs {st['tile_refs']} binned {st['triangles_binned']} max tile cycles {cyc.max()} sum cycles/592 {cyc.sum()/592:.0f}")
    r.close()

"""Tiny driver for an ncu launch list at small scale: C2 full frame, or one row band of C3 (usage: band_ncu.py c2 | c3 r0 r1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
cfg = sys.argv[1]
if cfg == "c2":
    sc, spec = scenes.scene_c2_terrain(voxel_dim=16, cube_size=32)
    W, H = 1920, 1080
else:
    sc, spec = scenes.scene_c3_instanced(voxel_dim=16, cube_size=32)
    W, H = 3840, 2160
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W, H)
if len(sys.argv) > 3:
    r.set_tile_rows(int(sys.argv[2]), int(sys.argv[3]))
for i in range(5):
    r.render_scene(sc, cam)
    r.resolve_device_only(2.0)
    r.synchronize()
print(r.stats())
r.close()

import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
sc, spec = scenes.scene_c3_instanced(voxel_dim=16, cube_size=32)
W,H=3840,2160
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W,H); r.render_scene(sc, cam, shade=False)
c = r.read_tile_counts().astype(np.int64)
print("tiles", c.size, "sum", c.sum(), "max", c.max(), "mean", c.mean().round(1), "p50", np.percentile(c,50), "p90", np.percentile(c,90), "p99", np.percentile(c,99))
top = np.sort(c.ravel())[::-1][:12]; print("top12", top)
print("row sums", c.sum(1))

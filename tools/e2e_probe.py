"""Where does the end-to-end frame time go on the host side? C3 at 4K, pipelined loop as in bench.py, with perf_counter
around every call; plus the pure host cost of enqueueing one frame when the GPU is idle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes

sc, spec = scenes.scene_c3_instanced(voxel_dim=128, cube_size=256)
W, H = 3840, 2160
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W, H)
bufs = [swr.RenderBuffer(W, H, pinned=True), swr.RenderBuffer(W, H, pinned=True)]
for _ in range(5):
    r.render_scene(sc, cam)
    r.blit_to_buffer(bufs[0])
N = 40
# (a) host cost of enqueueing a frame on an idle GPU
t_enq = []
for i in range(10):
    r.synchronize()
    t0 = time.perf_counter()
    r.render_scene(sc, cam)
    t_enq.append(time.perf_counter() - t0)
    r.synchronize()
print(f"enqueue one frame on an idle GPU (draw list + H2D + launches, no wait): {1e3 * np.median(t_enq):.3f} ms")
# (b) pipelined loop
seg = np.zeros(3)
prev = None
t_start = time.perf_counter()
for i in range(N):
    t0 = time.perf_counter()
    r.render_scene(sc, cam)
    t1 = time.perf_counter()
    tk = r.blit_to_buffer_async(bufs[i & 1])
    t2 = time.perf_counter()
    if prev is not None:
        r.wait_blit(prev)
    t3 = time.perf_counter()
    prev = tk
    seg += [t1 - t0, t2 - t1, t3 - t2]
r.wait_blit(prev)
tot = time.perf_counter() - t_start
st = r.stats()
print(f"pipelined: {1e3 * tot / N:.3f} ms/frame = render_scene {1e3 * seg[0] / N:.3f} (includes waiting for the previous frame) + blit_async {1e3 * seg[1] / N:.3f} "
      f"+ wait_blit {1e3 * seg[2] / N:.3f}; device phases {st['ms_setup_bin']:.3f}+{st['ms_raster']:.3f}+{st['ms_shade']:.3f}")
# (c) no read-back at all: render only, back to back
r.synchronize()
t0 = time.perf_counter()
for i in range(N):
    r.render_scene(sc, cam)
r.synchronize()
print(f"render only, back to back: {1e3 * (time.perf_counter() - t0) / N:.3f} ms/frame")
# (d) render + device-only resolve
t0 = time.perf_counter()
for i in range(N):
    r.render_scene(sc, cam)
    r.resolve_device_only(2.0)
r.synchronize()
print(f"render + device resolve, back to back: {1e3 * (time.perf_counter() - t0) / N:.3f} ms/frame")
r.close()

"""Experiment: two contexts (two streams, two sets of per-frame buffers) on ONE GPU, frames alternating between them, against
the single-context pipelined loop of bench.py. Does overlapping frame N+1's geometry with frame N's raster tail / shade pay?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes

sc, spec = scenes.scene_c3_instanced(voxel_dim=128, cube_size=256)
W, H = 3840, 2160
cam = swr.RenderCamera.from_spec(spec, W, H)
N = 60


def loop(rs, n):
    bufs = [swr.RenderBuffer(W, H, pinned=True) for _ in range(2 * len(rs))]
    pend = []
    t0 = time.perf_counter()
    for i in range(n):
        r = rs[i % len(rs)]
        r.render_scene(sc, cam)
        tk = r.blit_to_buffer_async(bufs[i % len(bufs)])
        pend.append((r, tk))
        if len(pend) > len(rs):
            rr, t = pend.pop(0)
            rr.wait_blit(t)
    for rr, t in pend:
        rr.wait_blit(t)
    return (time.perf_counter() - t0) / n * 1e3


one = [swr.Renderer(W, H)]
loop(one, 6)
print(f"one context, pipelined:  {loop(one, N):.3f} ms/frame")
two = [one[0], swr.Renderer(W, H)]
loop(two, 8)
print(f"two contexts, alternating: {loop(two, N):.3f} ms/frame")
print(f"two contexts, alternating: {loop(two, N):.3f} ms/frame")
print(f"one context, pipelined:  {loop(one, N):.3f} ms/frame")
for r in two:
    r.close()

"""Emulate every rank of an N-way sort-first split on ONE GPU: cost-balanced row bands from a probe frame, then each band
rendered on its own and timed (phase timers + device events around render+resolve). max over bands ~ the N-GPU frame
minus the strip gather.   usage: band_probe.py c3|c4 [N=8] [frames=8]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
from swraster_viewer_b200.multigpu import balanced_row_ranges

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
if cfg == "c4":
    sc, spec = scenes.scene_c4_micro(5001, voxel_dim=16, cube_size=32)
else:
    sc, spec = scenes.scene_c3_instanced(voxel_dim=16, cube_size=32)
W, H = 3840, 2160
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W, H)
stream = torch.cuda.ExternalStream(r.cuda_stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    r.render_scene(sc, cam, shade=False)
cyc = r.read_tile_costs()[1].astype(np.int64)
bands = [(0, r.tiles_y)] + balanced_row_ranges(cyc, N)
worst = 0.0
for rows in bands:
    r.set_tile_rows(*rows)
    tot = 0.0
    ph = np.zeros(3)
    for i in range(frames + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xFF)
            e0.record(stream)
        r.render_scene(sc, cam)
        r.resolve_device_only(2.0)
        e1.record(stream)
        r.synchronize()
        if i >= 3:
            st = r.stats()
            tot += e0.elapsed_time(e1)
            ph += [st["ms_setup_bin"], st["ms_raster"], st["ms_shade"]]
    st = r.stats()
    ms = tot / frames
    if rows != bands[0]:
        worst = max(worst, ms)
    print(f"[{cfg}] rows {rows}: frame {ms:.3f} ms | setup+bin {ph[0] / frames:.3f} raster {ph[1] / frames:.3f} shade {ph[2] / frames:.3f} | draws {r.num_draws} "
          f"T={st['triangles_submitted']} culled_clusters={st['clusters_culled']} binned={st['triangles_binned']} R={st['tile_refs']}", flush=True)
print(f"[{cfg}] N={N}: worst band {worst:.3f} ms -> {1000.0 / worst:.0f} frames/s before the gather; full frame = first line")
del stream, flush
r.close()

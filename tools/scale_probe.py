"""Per-rank breakdown of a sort-first frame (launch with torchrun): where do the milliseconds of an N-GPU frame go?
For every rank: rows owned, draws/triangles submitted, clusters culled, the three phase timers of swr_get_stats, the
wall time of render_scene (host draw list + launches + sync), resolve and the strip gather.
usage: scale_probe.py c3|c4 [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
from swraster_viewer_b200.multigpu import balanced_row_ranges, gather_strips, device_tensor

cfg = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = f"cuda:{local}"
W, H = 3840, 2160
if cfg == "c4":
    sc, spec = scenes.scene_c4_micro(5001, voxel_dim=128, cube_size=256)
else:
    sc, spec = scenes.scene_c3_instanced(voxel_dim=128, cube_size=256)
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W, H, device=local)
stream = torch.cuda.ExternalStream(r.cuda_stream(), device=local)
ranges = [(0, r.tiles_y)]
if world > 1:
    for _ in range(3):
        r.render_scene(sc, cam, shade=False)
    cyc = torch.from_numpy(r.read_tile_costs()[1].astype(np.int64)).to(dev)
    dist.broadcast(cyc, src=0)
    ranges = balanced_row_ranges(cyc.cpu().numpy(), world)
    r.set_tile_rows(*ranges[rank])
pix = device_tensor(r.device_pixels_ptr(), W * H * 4, torch.int32, dev).view(H, W)


def sync():
    stream.synchronize()


acc = {"render_wall": 0.0, "resolve_wall": 0.0, "gather_wall": 0.0, "ms_setup_bin": 0.0, "ms_raster": 0.0, "ms_shade": 0.0}
for i in range(steps + 3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r.render_scene(sc, cam)
    r.synchronize()
    t1 = time.perf_counter()
    r.resolve_device_only(2.0)
    sync()
    t2 = time.perf_counter()
    if world > 1:
        with torch.cuda.stream(stream):
            gather_strips(pix, ranges, H, dst=0)
        sync()
    t3 = time.perf_counter()
    if i >= 3:
        s = r.stats()
        acc["render_wall"] += (t1 - t0) * 1e3
        acc["resolve_wall"] += (t2 - t1) * 1e3
        acc["gather_wall"] += (t3 - t2) * 1e3
        for k in ("ms_setup_bin", "ms_raster", "ms_shade"):
            acc[k] += s[k]
s = r.stats()
out = {"rank": rank, "rows": ranges[rank] if world > 1 else ranges[0], "draws": r.num_draws, "T": s["triangles_submitted"], "binned": s["triangles_binned"],
       "R": s["tile_refs"], "culled": s["clusters_culled"]}
out.update({k: round(v / steps, 3) for k, v in acc.items()})
for k in range(world):
    if k == rank:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
del pix, stream
r.close()
if world > 1:
    dist.destroy_process_group()

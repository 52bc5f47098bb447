"""Multi-GPU runs of the BASELINE configs that are not the bench line: C4 (50 M micro-triangles, 3840x2160, sort-first)
and C5 (8 x 25 M-triangle shards, 7680x4320, sort-last with u64-min depth composite). Launch with torchrun.
Each run is timed (device events, max over ranks) and then VERIFIED: the multi-GPU image must equal rank 0's own
single-GPU render of the whole scene, pixel for pixel. With --oracle (use a reduced nverts: the serial CPU oracle holds
every 288-byte packet of the frame in memory) the multi-GPU visibility buffer (depth bits, seq) and image are ALSO compared
with the CPU oracle's.   usage: config_runs.py c4|c5 [steps] [nverts] [--oracle]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
from swraster_viewer_b200.multigpu import balanced_row_ranges, PeerAssembly, device_tensor, sort_last_frame

with_oracle = "--oracle" in sys.argv
sys.argv = [a for a in sys.argv if a != "--oracle"]
cfg = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = f"cuda:{local}"
t0 = time.time()
if cfg == "c4":
    W, H = 3840, 2160
    sc, spec = scenes.scene_c4_micro(int(sys.argv[3]) if len(sys.argv) > 3 else 5001, voxel_dim=128, cube_size=256)
else:
    W, H = 7680, 4320
    sc, spec = scenes.scene_c5_shards(8, int(sys.argv[3]) if len(sys.argv) > 3 else 3537, voxel_dim=128, cube_size=256)
gen_s = time.time() - t0
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W, H, device=local)
stream = torch.cuda.ExternalStream(r.cuda_stream(), device=local)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if cfg == "c4":
    ranges = [(0, r.tiles_y)]
    if world > 1:
        for _ in range(2):
            r.render_scene(sc, cam, shade=False)
        cyc = torch.from_numpy(r.read_tile_costs()[1].astype(np.int64)).to(dev)
        dist.broadcast(cyc, src=0)
        ranges = balanced_row_ranges(cyc.cpu().numpy(), world)
        r.set_tile_rows(*ranges[rank])
    pix = None
    pa = PeerAssembly(r, dst=0) if world > 1 else None

    def frame():
        global pix
        r.render_scene(sc, cam)
        pix = device_tensor(r.device_pixels_ptr(), W * H * 4, torch.int32, dev).view(H, W)
        if world > 1:
            pa.frame(2.0)  # the other ranks' resolve kernels store their rows into rank 0's buffer (NVLink peer memory)
            pa.release()
        else:
            r.resolve_device_only(2.0)
else:
    def frame():
        global pix
        if world > 1:
            pix = sort_last_frame(r, sc, cam, rank, world, stream)
        else:
            r.render_scene(sc, cam)
            r.resolve_device_only(2.0)
            pix = device_tensor(r.device_pixels_ptr(), W * H * 4, torch.int32, dev)

for _ in range(3):
    frame()
    r.synchronize()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    frame()
e1.record(stream)
stream.synchronize()
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
st = r.stats()
result = pix.clone().cpu().numpy().reshape(-1).view(np.uint32) if rank == 0 else None
vis = None
if with_oracle:
    # the frame's visibility buffer as the ranks hold it: depth is global after the key composite (sort-last) / per band
    # (sort-first); seq is known to the rank that owns the winner (others read FOREIGN / uncovered): unsigned MIN over ranks
    d, sq, _, _ = r.read_visbuffer()
    dt = torch.from_numpy(d.astype(np.int64)).to(dev)
    qt = torch.from_numpy(sq.astype(np.int64)).to(dev)
    if world > 1:
        if cfg == "c4":  # sort-first: rows outside my band hold nothing of this frame
            H64 = (H + 63) // 64
            rows = torch.zeros(H, dtype=torch.bool, device=dev)
            r0, r1 = ranges[rank]
            rows[r0 * 64:min(r1 * 64, H)] = True
            mask = rows.repeat_interleave(W)
            dt = torch.where(mask, dt, torch.full_like(dt, 1 << 40))
            qt = torch.where(mask, qt, torch.full_like(qt, 1 << 40))
        dist.all_reduce(dt, op=dist.ReduceOp.MIN)
        dist.all_reduce(qt, op=dist.ReduceOp.MIN)
    vis = (dt.cpu().numpy().astype(np.uint32), qt.cpu().numpy().astype(np.uint32))
barrier()
if rank == 0:
    # verification: the whole scene on this one GPU
    r.set_tile_rows(0, r.tiles_y)
    r.render_scene(sc, cam)
    buf = swr.RenderBuffer(W, H)
    r.blit_to_buffer(buf)
    same = bool(np.array_equal(buf.pixels, result))
    T = sc.total_triangles
    line = {"config": cfg, "n_gpus": world, "width": W, "height": H, "scene_triangles": T, "ms_per_frame": float(ms[0]), "frames_per_sec": 1e3 / float(ms[0]),
            "mtriangles_per_sec": T / float(ms[0]) / 1e3, "mode": "sort-first (balanced tile-row bands, peer-store frame assembly)" if cfg == "c4" else "sort-last (u64-min key composite, bary/pixel sum over NCCL)",
            "equals_single_gpu_render": same, "scene_build_s": round(gen_s, 1), "rank0_stats": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in st.items()}}
    if with_oracle:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        t0 = time.time()
        o = orc.Oracle(W, H)
        ref = o.render(sc, cam.abi, nthreads=1)
        opix = o.resolve(2.0)
        seq = np.where(vis[1] == 0xFFFFFFFE, 0xFFFFFFFF, vis[1]).astype(np.uint32)
        rgb = lambda p: np.stack([(p >> 24) & 255, (p >> 16) & 255, (p >> 8) & 255], -1).astype(np.int32)
        err = np.abs(rgb(result) - rgb(opix))
        line["oracle"] = {"depth_bits_equal": bool(np.array_equal(vis[0], ref["depth"])), "seq_equal": bool(np.array_equal(seq, ref["seq"])),
                          "pixels_differing_in_seq": int(np.count_nonzero(seq != ref["seq"])), "rgba8_max_err": int(err.max()), "rgba8_mean_err": float(err.mean()),
                          "oracle_seconds": round(time.time() - t0, 1)}
    print(json.dumps(line), flush=True)
    assert same, "multi-GPU image differs from the single-GPU render"
    if with_oracle:
        assert line["oracle"]["depth_bits_equal"] and line["oracle"]["seq_equal"] and line["oracle"]["rgba8_max_err"] <= 1, line["oracle"]
del pix, stream
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
r.close()

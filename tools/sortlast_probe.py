"""torchrun probe of the sort-last path (C5-like scene): NCCL u64-min key composite + bary/pixel sum."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import scenes
from swraster_viewer_b200.multigpu import sort_last_frame

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nverts = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
W, H = (7680, 4320) if len(sys.argv) < 3 else (int(sys.argv[2]), int(sys.argv[3]))
sc, spec = scenes.scene_c5_shards(8, nverts, voxel_dim=64, cube_size=128)
cam = swr.RenderCamera.from_spec(spec, W, H)
r = swr.Renderer(W, H, device=local)
stream = torch.cuda.ExternalStream(r.cuda_stream(), device=local)
for i in range(4):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    pix = sort_last_frame(r, sc, cam, rank, world, stream)
    stream.synchronize(); torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) * 1e3
if rank == 0:
    st = r.stats()
    print(f"sort-last x{world}: {sc.total_triangles} tris, {W}x{H}: {dt:.2f} ms/frame; rank0 stats {st}")
    img = pix.cpu().numpy().view(np.uint32)
    print("nonzero pixels", np.count_nonzero(img), "of", img.size)
    if W * H <= 1 << 21:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        import oracle as orc
        o = orc.Oracle(W, H); o.render(sc, cam.abi, nthreads=8); ref = o.resolve(2.0)
        print("max RGBA8 err vs oracle:", np.abs((img.view(np.uint8).astype(int) - ref.view(np.uint8).astype(int))).max())
del pix, stream
dist.barrier(); dist.destroy_process_group(); r.close()

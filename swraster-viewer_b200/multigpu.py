"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md 8(e).

sort-first: every rank holds the whole scene, owns a contiguous range of 64-pixel tile rows
(Renderer.set_tile_rows), and the finished RGBA8 strips are gathered to rank 0. Tiles are independent
after binning (renderer.rs:216-218), so there is no data-path collective besides that gather.

sort-last: every rank rasterises its own draws at full resolution into the 64-bit key buffer; the keys are
min-reduced across ranks (depth in the high word, inverted global seq in the low word, so the minimum IS the
reference's serial `<=` rule, tilerasterizer.rs:516) and the winner is shaded.
"""
import numpy as np


def tile_row_ranges(tiles_y, world):
    """Contiguous, balanced [begin, end) tile-row ranges, one per rank."""
    base, rem = divmod(tiles_y, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < rem else 0)
        out.append((r, r + n))
        r += n
    return out


def balanced_row_ranges(tile_cost, world, pixel_cost=None):
    """Contiguous tile-row bands with (nearly) equal cost. tile_cost: (tiles_y, tiles_x) per-tile raster cost from a
    full probe frame (Renderer.read_tile_costs: measured raster cycles, or ref counts) — identical on every rank up to
    timing noise, so rank 0's split is broadcast by the caller when cycles are used. Row cost = raster cost + a
    per-tile shading share (`pixel_cost`, default: 0.9 x mean raster cost, the measured shade:raster ratio on C3)."""
    tiles_y, tiles_x = tile_cost.shape
    tc = tile_cost.astype(np.float64)
    if pixel_cost is None:
        pixel_cost = 0.9 * tc.mean() + 1.0
    cost = tc.sum(axis=1) + pixel_cost * tiles_x
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        r = int(np.searchsorted(cum, target))
        r = r if abs(cum[min(r, tiles_y)] - target) <= abs(cum[max(r - 1, 0)] - target) else r - 1
        cuts.append(min(max(r, cuts[-1] + (1 if tiles_y - cuts[-1] > world - k else 0)), tiles_y - (world - k)))
    cuts.append(tiles_y)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def rebalance_row_ranges(ranges, band_ms, row_cost):
    """One feedback step of the sort-first split: `band_ms[k]` is what rank k's band (ranges[k]) really took per frame (all
    phases: the probe's raster cycles know nothing of a band's set-up or shading share); the rows of a band keep their relative
    `row_cost` (tiles_y values) but the band as a whole is re-weighted to its measured time, and the cuts are laid again so that
    every band gets the same share. Returns the new ranges (contiguous, every band at least one row)."""
    world, tiles_y = len(ranges), len(row_cost)
    w = np.asarray(row_cost, np.float64) + 1e-9
    dens = np.zeros(tiles_y)
    for (r0, r1), ms in zip(ranges, band_ms):
        if r1 > r0:
            dens[r0:r1] = w[r0:r1] / w[r0:r1].sum() * max(float(ms), 1e-6)
    cum = np.concatenate([[0.0], np.cumsum(dens)])
    cuts = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        r = int(np.searchsorted(cum, target))
        r = r if abs(cum[min(r, tiles_y)] - target) <= abs(cum[max(r - 1, 0)] - target) else r - 1
        cuts.append(min(max(r, cuts[-1] + 1), tiles_y - (world - k)))
    cuts.append(tiles_y)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def gather_strips(image, ranges, height, dst=0):
    """image: (H, W) int32 torch tensor whose rows [64*r0, 64*r1) are valid on this rank.
    Returns the assembled image on `dst` (in place in `image`), None elsewhere. Uses send/recv so strips of
    unequal height need no padding; on NCCL this is one grouped NVLink transfer per peer."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return image
    ops = []
    if rank == dst:
        for src in range(world):
            if src == dst:
                continue
            y0, y1 = ranges[src][0] * 64, min(ranges[src][1] * 64, height)
            if y1 > y0:
                ops.append(dist.P2POp(dist.irecv, image[y0:y1], src))
    else:
        y0, y1 = ranges[rank][0] * 64, min(ranges[rank][1] * 64, height)
        if y1 > y0:
            ops.append(dist.P2POp(dist.isend, image[y0:y1], dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return image if rank == dst else None


class PeerAssembly:
    """Sort-first frame assembly without a collective on the data path: contributing ranks' resolve kernels store their
    rows straight into the assembling rank's pixel buffer over NVLink peer memory and signal it (include/swr.h,
    swr_peer_*). torch.distributed only carries the 64-byte handle once. Per frame:
        pa.frame(exposure)            # every rank, after render_scene; returns once the work is enqueued
        ... rank dst consumes renderer.device_pixels_ptr() on the renderer's stream ...
        pa.release()                  # rank dst: the buffer may be overwritten by the next frame
    """

    def __init__(self, renderer, dst=0):
        import torch
        import torch.distributed as dist
        self.r, self.dst = renderer, dst
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.f = 0
        obj = [renderer.peer_export() if self.rank == dst else None]
        dist.broadcast_object_list(obj, src=dst)
        if self.rank != dst:
            renderer.peer_open(obj[0])
        dist.barrier()

    def frame(self, exposure):
        self.f += 1
        if self.rank == self.dst:
            self.r.resolve_device_only(exposure)
            self.r.peer_collect(self.f, self.world - 1)
        else:
            self.r.resolve_peer(exposure, self.f)

    def release(self):
        if self.rank == self.dst:
            self.r.peer_release(self.f)


class SharedHostFrame:
    """The caller's W*H u32 frame as ONE host buffer shared by the rank processes of a box: a /dev/shm segment every rank
    maps and page-locks (cudaHostRegister). Sort-first ranks then read their own rows back over their own PCIe links,
    in parallel, straight into the final frame (Renderer.blit_to_buffer copies only the rows a renderer owns), instead of
    funnelling 4*W*H bytes through rank 0's link. Behind the pixels: one u64 arrival counter per rank
    (arrive / wait_all) so the consumer knows when frame f is complete without a collective."""

    MAX_RANKS = 64

    def __init__(self, width, height, tag):
        import mmap
        import os
        import torch
        import torch.distributed as dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.nbytes_pix = width * height * 4
        n = self.nbytes_pix + self.MAX_RANKS * 8
        self.path = f"/dev/shm/{tag}"
        if self.rank == 0:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, n)
        if dist.is_initialized():
            dist.barrier()
        if self.rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self.mm = mmap.mmap(fd, n)
        os.close(fd)
        buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.pixels = buf[:self.nbytes_pix].view(np.uint32)
        self.flags = buf[self.nbytes_pix:].view(np.uint64)
        self._rt = torch.cuda.cudart()
        rc = self._rt.cudaHostRegister(self.pixels.ctypes.data, self.nbytes_pix, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {rc}")
        self._registered = True
        if dist.is_initialized():
            dist.barrier()

    def reset(self):
        self.flags[self.rank] = 0

    def arrive(self, rank, frame):
        self.flags[rank] = frame

    def wait_all(self, world, frame, timeout_s=10.0):
        import time
        t0 = time.perf_counter()
        while int(self.flags[:world].min()) < frame:
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError(f"SharedHostFrame.wait_all timed out waiting for frame {frame}: arrival flags {self.flags[:world].tolist()}")

    def close(self, unlink=False):
        import os
        if self._registered:
            self._rt.cudaHostUnregister(self.pixels.ctypes.data)
            self._registered = False
        self.pixels = self.flags = None
        try:
            self.mm.close()
        except BufferError:
            pass
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def composite_keys_min(keys):
    """keys: int64 torch tensor holding the UNSIGNED 64-bit visibility keys of this rank (empty = all ones).
    All-reduce with unsigned MIN: flip the sign bit so signed order == unsigned order, reduce, flip back."""
    import torch
    import torch.distributed as dist
    sign = torch.tensor(-(2 ** 63), dtype=torch.int64, device=keys.device)
    keys.bitwise_xor_(sign)
    dist.all_reduce(keys, op=dist.ReduceOp.MIN)
    keys.bitwise_xor_(sign)
    return keys


def device_tensor(ptr, nbytes, dtype, device):
    """Zero-copy torch view of a device buffer owned by the renderer (swr_device_pixels / swr_device_keys)."""
    import torch

    class _Arr:
        pass
    itemsize = torch.empty((), dtype=dtype).element_size()
    a = _Arr()
    typestr = {torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    a.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device=device)


def sort_last_frame(r, scene, camera, rank, world, stream, exposure=2.0):
    """One sort-last frame on this rank: draws are dealt round-robin (draw index % world), keys are
    min-reduced, barycentrics sum-reduced, every rank shades the pixels whose winner it owns, RGBA8 is sum-reduced.
    Every rank ends up with the full image in its device pixel buffer. `stream`: torch ExternalStream of r."""
    import torch
    import torch.distributed as dist
    dev = f"cuda:{torch.cuda.current_device()}"
    r.render_scene(scene, camera, shade=False, shard=rank, nshards=world)
    r.keys_to_global()
    kptr, kbytes = r.device_keys_ptr()
    keys = device_tensor(kptr, kbytes, torch.int64, dev)
    with torch.cuda.stream(stream):
        composite_keys_min(keys)
    r.keys_localize()
    bary = device_tensor(r.device_bary_ptr(), r.width * r.height * 8, torch.int32, dev).view(torch.float32)
    with torch.cuda.stream(stream):
        dist.all_reduce(bary, op=dist.ReduceOp.SUM)
    rows = tile_row_ranges(r.tiles_y, world)[rank]
    r.shade_composited(camera, rows[0], rows[1])
    r.resolve_device_only(exposure)
    pix = device_tensor(r.device_pixels_ptr(), r.width * r.height * 4, torch.int32, dev)
    with torch.cuda.stream(stream):
        dist.all_reduce(pix, op=dist.ReduceOp.SUM)
    return pix

"""Sort-last (shard by primitive + depth composite) emulated on ONE GPU: two contexts render disjoint
draws, the cross-rank collectives (u64 MIN on keys, SUM on barycentrics, SUM on RGBA8) are done with plain torch ops.
The composited visibility must equal the oracle's full frame bit for bit, its RGBA8 within 1 LSB. (The NCCL/gloo side is covered by
tests/test_host_mirror.py::test_sort_last_key_composite_gloo_world2.)"""
import numpy as np
import pytest

import swraster_viewer_b200 as swr
from swraster_viewer_b200.multigpu import device_tensor, tile_row_ranges
from helpers import small_configs, render_oracle, rgba_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [4, 2])
def test_sort_last_two_shards_equal_full_frame(cfg):
    import torch
    name, scene, spec, W, H = small_configs()[cfg]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    world = 2
    rs = [swr.Renderer(W, H) for _ in range(world)]
    dev = "cuda:0"
    for k, r in enumerate(rs):
        r.render_scene(scene, cam, shade=False, shard=k, nshards=world)
        r.keys_to_global()
        r.synchronize()
    sign = torch.tensor(-(2 ** 63), dtype=torch.int64, device=dev)
    keys = [device_tensor(*r.device_keys_ptr(), torch.int64, dev) for r in rs]
    kmin = torch.minimum(keys[0] ^ sign, keys[1] ^ sign) ^ sign  # unsigned 64-bit min
    for k in keys:
        k.copy_(kmin)
    torch.cuda.synchronize()
    for r in rs:
        r.keys_localize()
        r.synchronize()
    barys = [device_tensor(r.device_bary_ptr(), W * H * 8, torch.int32, dev).view(torch.float32) for r in rs]
    bsum = barys[0] + barys[1]
    for b in barys:
        b.copy_(bsum)
    torch.cuda.synchronize()
    rows = tile_row_ranges(rs[0].tiles_y, world)
    pix, seqs, depths = [], [], []
    for k, r in enumerate(rs):
        r.shade_composited(cam, *rows[k])
        r.resolve_device_only(2.0)
        r.synchronize()
        pix.append(device_tensor(r.device_pixels_ptr(), W * H * 4, torch.int32, dev).clone())
        d, s, _, _ = r.read_visbuffer()
        seqs.append(s)
        depths.append(d)
    torch.cuda.synchronize()
    total = (pix[0] + pix[1]).cpu().numpy().view(np.uint32)
    o = render_oracle(scene, cam, W, H)
    # every pixel is shaded by exactly one rank
    owned = [(p.cpu().numpy().view(np.uint32) != 0) for p in pix]
    assert np.all(owned[0] ^ owned[1])
    FOREIGN = 0xFFFFFFFE
    seq = np.where(seqs[0] == FOREIGN, seqs[1], seqs[0])
    assert np.array_equal(seq, o["seq"])
    assert np.array_equal(depths[0], o["depth"]) and np.array_equal(depths[1], o["depth"])
    err = np.abs(rgba_bytes(total) - rgba_bytes(o["pixels"]))
    # visibility (seq, depth) is bit-exact above; colour is within the +-1 LSB budget (value-domain FMAs in the shader)
    assert err.max() <= 1, f"{name}: composited RGBA8 differs (max {err.max()})"
    for r in rs:
        r.close()

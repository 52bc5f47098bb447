"""Fixed-exposure frames (include/swr.h swr_set_fixed_exposure): shading packs RGBA8 itself (renderer.rs:293-355 folded into
the shading pass) and produces the tile metering values (tilerasterizer.rs:103-106) in the same kernel. The bar: exactly the
pixels and metering values of the HDR path (k_shade -> k_luminance -> k_resolve), which is the one compared with the oracle
everywhere else; and the host mirror falls back to HDR by itself when update_auto_exposure moves the exposure."""
import ctypes as C

import numpy as np
import pytest

import swraster_viewer_b200 as swr
from helpers import small_configs, render_oracle, rgba_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def configs():
    return small_configs()


@pytest.mark.parametrize("idx", [0, 2, 5])
def test_fixed_exposure_frame_equals_hdr_path(configs, idx):
    name, scene, spec, W, H = configs[idx]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    buf = swr.RenderBuffer(W, H)
    r.render_scene(scene, cam)  # the meter is idle: shaded straight to RGBA8 with the exposure held (2.0)
    r.blit_to_buffer(buf)
    fused_px = buf.pixels.copy()
    fused_lum = r.read_tile_luminance()
    launches_fused = r.launch_count
    hdr = r.read_color()        # asks for HDR colour: the mirror shades the same visibility buffer again, to HDR
    assert r.launch_count > launches_fused, "read_color of a fixed-exposure frame must re-shade"
    r.blit_to_buffer(buf)       # k_resolve of that HDR colour
    assert np.array_equal(fused_px, buf.pixels), f"{name}: fixed-exposure RGBA8 differs from the HDR path's"
    assert np.array_equal(fused_lum.view(np.uint32), r.read_tile_luminance().view(np.uint32)), f"{name}: metering values differ"
    assert np.isfinite(hdr).all()
    o = render_oracle(scene, cam, W, H)
    assert np.abs(rgba_bytes(fused_px) - rgba_bytes(o["pixels"])).max() <= 1
    r.close()


def test_fixed_exposure_pipelined_lanes(configs):
    """Two lanes, asynchronous read-back: every frame of a fixed-exposure sequence equals the synchronous HDR-path frame."""
    name, scene, spec, W, H = configs[2]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    ref = swr.Renderer(W, H)
    rb = swr.RenderBuffer(W, H)
    ref.render_scene(scene, cam)
    ref.read_color()
    ref.blit_to_buffer(rb)
    want = rb.pixels.copy()
    ref.close()
    r = swr.Renderer(W, H, lanes=2)
    bufs = [swr.RenderBuffer(W, H, pinned=True) for _ in range(3)]
    tickets = []
    for i in range(6):
        r.render_scene(scene, cam)
        tickets.append((r.blit_to_buffer_async(bufs[i % 3]), i % 3))
        if len(tickets) > 2:
            t, b = tickets.pop(0)
            r.wait_blit(t)
            assert np.array_equal(bufs[b].pixels, want), f"frame {i - 2} differs"
            bufs[b].pixels[:] = 0
    for t, b in tickets:
        r.wait_blit(t)
        assert np.array_equal(bufs[b].pixels, want)
    r.close()


def test_meter_moves_exposure_between_render_and_blit(configs):
    name, scene, spec, W, H = configs[0]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    buf = swr.RenderBuffer(W, H)
    r.render_scene(scene, cam)
    r.update_auto_exposure(0.5)  # the exposure leaves 2.0 after the frame was shaded with 2.0
    e = r.auto_exposure
    assert e != pytest.approx(2.0)
    r.blit_to_buffer(buf)        # the mirror shades again (HDR) and resolves with the new exposure
    import oracle as orc
    o = orc.Oracle(W, H)
    o.render(scene, cam.abi)
    assert np.abs(rgba_bytes(buf.pixels) - rgba_bytes(o.resolve(e))).max() <= 1
    r.close()


def test_c_abi_refuses_another_exposure(configs):
    name, scene, spec, W, H = configs[0]
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H)
    r.render_scene(scene, cam)
    core = r.core
    out = np.zeros(W * H, np.uint32)
    assert core.swr_resolve(r.ctx, C.c_float(1.5), out.ctypes.data) != 0, "a fixed-exposure frame cannot be resolved with another exposure"
    assert b"fixed exposure" in core.swr_last_error(r.ctx)
    assert core.swr_resolve(r.ctx, C.c_float(2.0), out.ctypes.data) == 0
    assert core.swr_set_fixed_exposure(r.ctx, C.c_float(-1.0)) != 0
    assert core.swr_set_fixed_exposure(r.ctx, C.c_float(0.0)) == 0
    assert core.swr_shade(r.ctx, C.byref(cam.abi)) == 0  # same visibility buffer, now HDR
    out2 = np.zeros(W * H, np.uint32)
    assert core.swr_resolve(r.ctx, C.c_float(1.5), out2.ctypes.data) == 0
    assert (out2 != out).any()
    r.close()

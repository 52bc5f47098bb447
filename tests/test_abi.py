"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/swr.h declares, and its struct layouts match the ctypes mirror. No compute calls here."""
import ctypes as C
import os
import re

import pytest

import swraster_viewer_b200 as swr
from swraster_viewer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    core, host = swr.load_libraries()
    header = open(os.path.join(ROOT, "include", "swr.h")).read()
    declared = set(re.findall(r"\b(swr_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert hasattr(core, name), f"libswr_b200.so does not export {name}"
    assert core.swr_abi_version() == 1


def test_gltf_and_host_header_symbols_are_exported_by_the_host_library():
    _, host = swr.load_libraries()
    total = set()
    for hdr, must in (("swr_gltf.h", {"swrh_gltf_load", "swrh_gltf_scene", "swrh_gltf_free", "swrh_build_mip_chain", "swrh_decode_png", "swrh_env_bake",
                                      "swrh_compute_sun_visibility"}),
                      ("swr_host.h", {"swrh_renderer_new", "swrh_render_scene", "swrh_blit_to_buffer", "swrh_update_auto_exposure", "swrh_camera_build",
                                      "swrh_build_draws"})):
        header = open(os.path.join(ROOT, "include", hdr)).read()
        declared = set(re.findall(r"\b(swrh_[a-z_0-9]+)\s*\(", header))
        assert must <= declared, must - declared
        for name in declared:
            assert hasattr(host, name), f"libswr_host.so does not export {name} ({hdr})"
        total |= declared
    # and nothing the library exports is left undeclared: every swrh_ symbol of the .so appears in one of the two headers
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "swraster-viewer_b200", "lib", "libswr_host.so")], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(swrh_[a-z_0-9]+)\b", out))
    assert exported == total, exported ^ total


def test_struct_sizes_match_library():
    core, _ = swr.load_libraries()
    for which, T in enumerate([abi.PrimitiveDesc, abi.MeshDesc, abi.NodeDesc, abi.TextureDesc, abi.MaterialDesc,
                               abi.VoxelGridDesc, abi.SceneDesc, abi.Camera, abi.Draw, abi.FrameStats]):
        assert core.swr_sizeof(which) == C.sizeof(T), (which, T.__name__)


def test_library_carries_sm100a_code_only():
    import subprocess
    lib = os.path.join(ROOT, "swraster-viewer_b200", "lib", "libswr_b200.so")
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "swraster-viewer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "orc_", "oracle/", "oracle.cpp", "oracle.py"):
                    assert needle not in text, (os.path.join(dirpath, f), needle)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    core, _ = swr.load_libraries()
    assert core.swr_create(64, 64, 0) is None
    assert b"no usable CUDA device" in core.swr_last_error(None) or b"CUDA" in core.swr_last_error(None)
    with pytest.raises(RuntimeError):
        swr.Renderer(64, 64)


def test_headers_are_c99_and_the_c_example_builds_and_fails_loudly_without_a_device(tmp_path):
    """include/*.h must be usable from plain C (the reference-side binding is an FFI, not C++): the example host program
    compiles with -std=c99 -pedantic -Werror against all three headers, loads a glTF file, and — on a box without a B200 —
    stops at Renderer::new with the 'no CPU fallback' message instead of producing an image some other way."""
    import subprocess
    from swraster_viewer_b200 import scenes
    lib = os.path.join(ROOT, "swraster-viewer_b200", "lib")
    exe = str(tmp_path / "render_gltf")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "render_gltf.c"), "-L", lib, "-lswr_host", "-lswr_b200", "-Wl,-rpath," + lib, "-lm", "-o", exe])
    sc, _ = scenes.scene_c1_sphere(segments=16, bands=8, voxel_dim=4, cube_size=8)
    scenes.export_gltf(sc, str(tmp_path / "s"))
    r = subprocess.run([exe, str(tmp_path / "s.gltf"), str(tmp_path / "out.ppm"), "128", "64"], capture_output=True, text=True)
    assert "1 primitives, 1 nodes, 1 materials" in r.stdout
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and os.path.getsize(tmp_path / "out.ppm") == len("P6\n128 64\n255\n") + 128 * 64 * 3
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr and not os.path.exists(tmp_path / "out.ppm")
    r = subprocess.run([exe, str(tmp_path / "missing.gltf"), str(tmp_path / "out.ppm")], capture_output=True, text=True)
    assert r.returncode == 1 and "Missing data" in r.stderr


def test_every_entry_point_survives_a_null_context():
    """The FFI caller may hold a failed (NULL) context: every entry point must return an error / NULL, not dereference it.
    Runs in a child process so that a regression shows up as a failure instead of killing the test run."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
import swraster_viewer_b200 as swr
from swraster_viewer_b200 import abi
core, _ = swr.load_libraries()
core.swr_resolve_async.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
core.swr_wait_pixels.argtypes = [C.c_void_p, C.c_int]
args = {"swr_set_rsqrt_table": (None, None, 10), "swr_set_tile_rows": (None, 0, 1), "swr_upload_scene": (None, None), "swr_share_scene": (None, None), "swr_render": (None, None, None, 0, 1),
        "swr_set_fixed_exposure": (None, 2.0), "swr_shade": (None, None), "swr_shade_composited": (None, None, 0, 0), "swr_resolve": (None, 2.0, None), "swr_resolve_async": (None, 2.0, None, None),
        "swr_wait_pixels": (None, 0), "swr_read_tile_luminance": (None, None), "swr_read_tile_costs": (None, None, None),
        "swr_read_visbuffer": (None, None, None, None, None), "swr_read_color": (None, None), "swr_get_stats": (None, None), "swr_peer_export": (None, None),
        "swr_peer_open": (None, None), "swr_peer_attach": (None, None), "swr_resolve_peer": (None, 2.0, 1), "swr_peer_collect": (None, 1, 1),
        "swr_peer_release": (None, 1), "swr_multi_tile_rows": (None, 0, None, None), "swr_multi_set_rsqrt_table": (None, None, 10), "swr_multi_upload_scene": (None, None),
        "swr_multi_render": (None, None, None, 0), "swr_multi_resolve": (None, 2.0, None), "swr_multi_read_tile_luminance": (None, None), "swr_multi_get_stats": (None, None),
        "swr_multi_context": (None, 0), "swr_bake_brdf_lut": (0, 0, None), "swr_bake_irradiance_sh4": (0, None, 4, 4, None),
        "swr_bake_prefilter_specular": (0, None, 4, 4, 8, None), "swr_bake_sun_visibility": (0, None, None)}
core.swr_multi_resolve.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
skip = {"swr_abi_version", "swr_last_error", "swr_create", "swr_sizeof", "swr_multi_create", "swr_multi_last_error", "swr_bake_last_error"}
for name in abi.EXPORTS:
    if name in skip:
        continue
    r = getattr(core, name)(*args.get(name, (None,)))
    ok = r is None or r == 0 if name in ("swr_destroy", "swr_device_pixels", "swr_device_keys", "swr_device_keys_bytes", "swr_device_bary", "swr_cuda_stream", "swr_launch_count", "swr_multi_destroy", "swr_multi_device_count",
                                     "swr_multi_context") else r < 0
    assert ok, (name, r)
print("NULL-SAFE", len(abi.EXPORTS) - len(skip))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "NULL-SAFE" in r.stdout, r.stderr[-2000:]


def test_host_entry_points_survive_null_arguments():
    """Same for the host mirror's C API (include/swr_host.h, include/swr_gltf.h): NULL handles and pointers are errors."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
import swraster_viewer_b200 as swr
_, h = swr.load_libraries()
h.swrh_auto_exposure.restype = C.c_float
h.swrh_renderer_ctx.restype = C.c_void_p
h.swrh_gltf_load.restype = C.c_void_p
h.swrh_gltf_scene.restype = C.c_void_p
h.swrh_gltf_texture_uri.restype = C.c_char_p
h.swrh_env_bake.restype = C.c_void_p
f = C.c_float
neg = {"swrh_camera_build": (None, None, f(1), f(64), f(64), f(10), None), "swrh_camera_build_rotated": (None, None, f(0), f(0), f(1), f(64), f(64), f(10), None),
       "swrh_set_reference_rsqrt": (None, 1), "swrh_reference_rsqrt_bits": (None,), "swrh_set_tile_rows": (None, 0, 1),
       "swrh_render_scene": (None, None, None, 1, 0, 1), "swrh_num_draws": (None,), "swrh_update_auto_exposure": (None, f(0)),
       "swrh_blit_to_buffer": (None, None, 64, 64), "swrh_blit_to_buffer_async": (None, None, 64, 64, None), "swrh_wait_blit": (None, 0),
       "swrh_build_draws": (None, None, None, 0, 0, 1), "swrh_build_draws_band": (None, None, None, 0, 0, 64, 64),
       "swrh_gltf_get_info": (None, None), "swrh_gltf_get_camera": (None, 0, None), "swrh_gltf_register_image": (None, None, 0, 0),
       "swrh_gltf_bake_sun_visibility": (None,), "swrh_compute_sun_visibility": (None, None), "swrh_env_get": (None, None, None)}
for name, a in neg.items():
    r = getattr(h, name)(*a)
    assert r < 0, (name, r)
for name, a in {"swrh_renderer_ctx": (None,), "swrh_gltf_load": (None, None), "swrh_gltf_scene": (None,), "swrh_gltf_texture_uri": (None, 0),
                "swrh_env_bake": (None, 4, 3, 4, 4, 1, f(1), f(1), f(1))}.items():
    assert getattr(h, name)(*a) is None, name
h.swrh_renderer_free(None); h.swrh_gltf_free(None); h.swrh_env_free(None)
assert h.swrh_auto_exposure(None) == 0.0
print("NULL-SAFE", len(neg))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "NULL-SAFE" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
